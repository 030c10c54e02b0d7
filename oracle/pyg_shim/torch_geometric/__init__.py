"""TEST INFRASTRUCTURE ONLY - stand-in for torch-geometric 2.3.1 (requirements.txt:20).

Only the six symbols the TwoWL path of the reference touches are restated:
nn.GCNConv, nn.GraphNorm (TwoWL/model/model.py:3), utils.negative_sampling,
utils.add_self_loops (TwoWL/operators/datasets.py:3), data.Data (datasets.py:2).
The real wheel is absent from this image and cannot be installed offline.
"""
__version__ = "2.3.1-shim"
