"""TEST INFRASTRUCTURE ONLY - ``Data`` as used at TwoWL/operators/datasets.py:165."""
import torch


class Data:
    """Attribute bag; ``num_nodes`` is inferred as max node id + 1 (PyG 2.3.1 behaviour
    when no ``x`` / explicit ``num_nodes`` is given)."""

    def __init__(self, edge_index=None, edge_attr=None, **kw):
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self._num_nodes = kw.pop("num_nodes", None)
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if self._num_nodes is not None:
            return self._num_nodes
        cands = [v for k, v in self.__dict__.items()
                 if isinstance(v, torch.Tensor) and "edge_index" in k and v.numel() > 0]
        return int(max(int(v.max()) for v in cands)) + 1 if cands else 0

    @num_nodes.setter
    def num_nodes(self, n):
        self._num_nodes = n
