"""Shim of the torch_geometric surface the reference touches."""
