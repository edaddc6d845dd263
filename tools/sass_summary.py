"""profiles/r2_sass_summary.md: per-kernel counts of the Blackwell-native SASS mnemonics in liblgcn_b200.so.
    cuobjdump -sass lanegcn-1_b200/liblgcn_b200.so > /tmp/sass.txt && python tools/sass_summary.py /tmp/sass.txt"""
import collections, re, subprocess, sys
txt = open(sys.argv[1]).read()
pat = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'ELECT', 'HMMA', 'FFMA']
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = subprocess.run(['c++filt', f.split('\n', 1)[0].strip()], capture_output=True, text=True).stdout.strip()
    ins = collections.Counter(re.findall(r'/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', f))
    print(name[:90], {p: ins[p] for p in pat if ins[p]}, sum(ins.values()))
