"""MessagePassing of PyG 1.6.3 restated from its published behaviour:
flow source_to_target; `foo_j` = kw['foo'][edge_index[0]], `foo_i` = kw['foo'][edge_index[1]];
messages reduced with torch_scatter.scatter at edge_index[1] along node_dim with dim_size = N."""
import inspect

import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        params = [p for p in inspect.signature(self.message).parameters]
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() >= 2:
                n = v.size(self.node_dim)
                break
        args = {}
        for name in params:
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs[name[:-2]]
                sel = edge_index[0] if name.endswith("_j") else edge_index[1]
                if torch.is_tensor(data):
                    n = data.size(self.node_dim)
                    data = data.index_select(self.node_dim, sel)
                args[name] = data
            else:
                args[name] = kwargs.get(name, None)
        msg = self.message(**args)
        out = self.aggregate(msg, edge_index[1], dim_size=n)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs
