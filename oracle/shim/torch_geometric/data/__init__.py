import torch


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Dataset(torch.utils.data.Dataset):
    pass


class Batch(Data):
    """Appendix A.10: concatenate on dim 0, 0-dim tensors become [B], add batch and ptr."""
    @staticmethod
    def from_data_list(items):
        out = Batch()
        keys = [k for k, v in items[0].__dict__.items() if torch.is_tensor(v)]
        for k in keys:
            vals = [getattr(d, k) for d in items]
            vals = [v.unsqueeze(0) if v.dim() == 0 else v for v in vals]
            setattr(out, k, torch.cat(vals, 0))
        n = torch.tensor([d.pos.size(0) for d in items])
        out.batch = torch.repeat_interleave(torch.arange(len(items)), n)
        out.ptr = torch.cat([n.new_zeros(1), n.cumsum(0)])
        return out
