import inspect

import torch


class MessagePassing(torch.nn.Module):
    """aggr='add', flow='source_to_target': out[dst] += message(x_j = x[src])."""

    def __init__(self, aggr='add', **kwargs):
        super().__init__()
        if aggr != 'add':
            raise NotImplementedError('stub supports aggr="add" only')

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        x = kwargs.get('x')
        wanted = inspect.signature(self.message).parameters
        call = {}
        for name in wanted:
            if name == 'x_j':
                call[name] = x[src]
            elif name == 'x_i':
                call[name] = x[dst]
            else:
                call[name] = kwargs[name]
        msg = self.message(**call)
        out = torch.zeros((x.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        out = out.index_add(0, dst, msg)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
