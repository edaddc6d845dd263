"""ORACLE — test infrastructure only (see oracle/lanegcn_oracle.py header).  Never imported by the product."""
