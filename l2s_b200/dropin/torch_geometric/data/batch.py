"""torch_geometric.data subset: Batch.from_data_list for edge lists (conventional_baselines...:175)."""
import types

import torch


class Data(types.SimpleNamespace):
    pass


class Batch:
    @staticmethod
    def from_data_list(datas):
        off, parts, batch = 0, [], []
        for k, d in enumerate(datas):
            parts.append(d.edge_index + off)
            batch.append(torch.full((d.num_nodes,), k, dtype=torch.long))
            off += d.num_nodes
        return Data(edge_index=torch.cat(parts, dim=-1), batch=torch.cat(batch), num_nodes=off)
