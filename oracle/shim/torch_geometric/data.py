"""Stand-ins for torch_geometric.data.{Data, InMemoryDataset, DataLoader}.

Semantics restated from the published PyG 1.x behaviour the reference relies on
(call sites: quantum/decoder_v2_4.py:161-174,205-206; classical/CGNNI.py:150-164,208-209):
  * Data            -- attribute bag with .to(device)
  * InMemoryDataset -- collate() keeps the list of Data objects
  * DataLoader      -- batches `batch_size` consecutive graphs block-diagonally: x and y are
                       concatenated along dim 0, edge_index of graph g is offset by the
                       cumulative node count (= x_g.size(0)) and concatenated along dim 1.
"""
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, y=None, **kw):
        self.x, self.edge_index, self.y = x, edge_index, y
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class InMemoryDataset(object):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.data_list = []

    def collate(self, data_list):
        self.data_list = list(data_list)
        return self.data_list, None

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, i):
        return self.data_list[i]


class DataLoader(object):
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        assert not shuffle, "oracle shim: reference only uses shuffle=False"
        self.dataset, self.batch_size = dataset, batch_size

    def __iter__(self):
        items = self.dataset.data_list
        for s in range(0, len(items), self.batch_size):
            chunk = items[s:s + self.batch_size]
            xs, ys, eis, off = [], [], [], 0
            for d in chunk:
                xs.append(d.x)
                ys.append(d.y)
                eis.append(d.edge_index + off)
                off += d.x.size(0)
            b = Data(x=torch.cat(xs, 0), y=torch.cat(ys, 0), edge_index=torch.cat(eis, 1))
            b.num_graphs = len(chunk)
            yield b

    def __len__(self):
        return (len(self.dataset.data_list) + self.batch_size - 1) // self.batch_size
