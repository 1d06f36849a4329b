import torch
from torch_geometric.data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        kw.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle,
                         collate_fn=Batch.from_data_list, **kw) if "batch_sampler" not in kw else \
            super().__init__(dataset, collate_fn=Batch.from_data_list, **kw)
