"""lanegcn_b200 — B200-native (sm_100a) LaneGCN forward graph path behind the reference's module API.

Layout
  lanegcn.py   drop-in for the reference ``lanegcn.py`` module API (Net, MapNet, A2M, M2M, M2A, A2A, Att,
               actor_gather, graph_gather, get_model, config)
  _C.py        ctypes binding of the C-ABI shared library (``csrc/`` -> ``liblgcn_b200.so``)
  synth.py     seeded synthetic Argoverse-shaped scenes (dataset is offline)
  shard.py     scene sharding across ranks + final result gather

The CUDA library is required: there is no CPU or PyTorch fallback for the hot path.
"""
__version__ = "0.1.0"
