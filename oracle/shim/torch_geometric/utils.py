import torch

def remove_self_loops(edge_index, edge_attr=None):
    m = edge_index[0] != edge_index[1]
    return edge_index[:, m], edge_attr

def add_self_loops(edge_index, edge_attr=None, num_nodes=None):
    loop = torch.arange(num_nodes, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], 1), edge_attr
