"""stgraph_b200 -- B200-native backend for STGraph's vertex-centric aggregation hot path.

Python surface mirrors ``stgraph`` (compiler decorator + IR, ``nn.pytorch`` layers,
graph containers); everything underneath runs in hand-written sm_100a CUDA kernels
reached through the C ABI of ``include/stgraph_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
