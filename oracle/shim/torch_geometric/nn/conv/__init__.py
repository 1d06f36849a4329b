"""MessagePassing with aggr='max', flow source_to_target."""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kw):
        super().__init__()
        assert aggr == "max" and flow == "source_to_target"
        self.aggr = aggr

    def reset_parameters(self):
        pass

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        args = {}
        dim_size = None
        for name in inspect.signature(self.message).parameters:
            if name == "edge_index_i":
                args[name] = i
            elif name == "edge_index_j":
                args[name] = j
            elif name.endswith("_j") or name.endswith("_i"):
                data = kwargs[name[:-2]]
                which = 0 if name.endswith("_j") else 1
                idx = j if which == 0 else i
                if isinstance(data, (tuple, list)):
                    if data[1] is not None:
                        dim_size = data[1].size(0)
                    data = data[which]
                args[name] = None if data is None else data.index_select(0, idx)
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        out = msg.new_zeros((dim_size,) + tuple(msg.shape[1:]))
        idx = i.view(-1, 1).expand_as(msg)
        return out.scatter_reduce(0, idx, msg, reduce="amax", include_self=False)
