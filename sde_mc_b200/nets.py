"""Control-variate function approximators and path datasets (API of /root/reference/sde_mc/nets.py).

Training stays in PyTorch autograd.  For inference, the BN-free `Mlp` used by the experiments is evaluated inside
the fused control-variate kernel (csrc/cv.cuh) on tensor cores; `mlp_layers()` exports its weights for that."""
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import Dataset


class ControlVariate(nn.Module):
    """Base class: `sequential` tells the appliers whether the net consumes whole paths (nets.py:7-20)."""

    def __init__(self, sequential, device):
        super().__init__()
        self.sequential = sequential
        self.device = device


class ZeroFunction(ControlVariate):
    """Maps everything to zeros of width output_dim (nets.py:23-36)."""

    def __init__(self, output_dim):
        super().__init__(sequential=False, device=None)
        self.output_dim = output_dim

    def forward(self, x):
        return torch.zeros((x.shape[0], self.output_dim), device=x.device)


class Mlp(ControlVariate):
    """[BatchNorm] Linear [BatchNorm] act ... Linear [final act]  (nets.py:39-93)."""

    def __init__(self, input_size, layer_sizes, output_size, activation=nn.ReLU, final_activation=None,
                 batch_norm=True, batch_norm_init=True, device='cpu'):
        assert len(layer_sizes) > 0, "At least one hidden layer required."
        super().__init__(sequential=False, device=device)
        self.num_layers = len(layer_sizes)
        widths = [input_size] + list(layer_sizes)
        mods = [nn.BatchNorm1d(input_size, device=device)] if batch_norm_init else []
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            mods.append(nn.Linear(fan_in, fan_out, device=device))
            if batch_norm:
                mods.append(nn.BatchNorm1d(fan_out, device=device))
            mods.append(activation())
        mods.append(nn.Linear(widths[-1], output_size, device=device))
        if final_activation is not None:
            mods.append(final_activation())
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)

    def mlp_layers(self):
        """[(weight, bias)] if the net is a plain Linear/ReLU stack the fused kernel can evaluate, else None."""
        layers = []
        mods = list(self.net)
        for i, mod in enumerate(mods):
            if isinstance(mod, nn.Linear):
                layers.append((mod.weight, mod.bias))
            elif isinstance(mod, nn.ReLU):
                if i == len(mods) - 1:
                    return None
            else:
                return None
        return layers


class Lstm(ControlVariate):
    """LSTM over whole paths + linear head (nets.py:96-128)."""

    def __init__(self, in_dim, hidden_dim, out_dim, device='cpu'):
        super().__init__(sequential=True, device=device)
        self.in_dim, self.hidden_dim, self.out_dim = in_dim, hidden_dim, out_dim
        self.lstm = nn.LSTM(in_dim, hidden_dim, batch_first=True).to(device)
        self.lin = nn.Linear(hidden_dim, out_dim).to(device)

    def init_hidden(self, bs):
        zeros = torch.zeros((1, bs, self.hidden_dim), device=self.device)
        return zeros, zeros.clone()

    def forward(self, x):
        out, _ = self.lstm(x, self.init_hidden(x.shape[0]))
        return self.lin(out)


class Gru(ControlVariate):
    """GRU over whole paths + linear head (nets.py:131-162)."""

    def __init__(self, in_dim, hidden_dim, out_dim, device='cpu'):
        super().__init__(sequential=True, device=device)
        self.in_dim, self.hidden_dim, self.out_dim = in_dim, hidden_dim, out_dim
        self.gru = nn.GRU(in_dim, hidden_dim, batch_first=True).to(device)
        self.lin = nn.Linear(hidden_dim, out_dim).to(device)

    def init_hidden(self, bs):
        return torch.zeros((1, bs, self.hidden_dim), device=self.device)

    def forward(self, x):
        out, _ = self.gru(x, self.init_hidden(x.shape[0]))
        return self.lin(out)


class NormalPathData(Dataset):
    """(paths without the terminal state, increments) -> payoff (nets.py:165-175)."""

    def __init__(self, paths, payoffs, normals):
        self.paths = paths[:, :-1]
        self.payoffs = payoffs
        self.normals = normals

    def __len__(self):
        return len(self.payoffs)

    def __getitem__(self, idx):
        return (self.paths[idx], self.normals[idx]), self.payoffs[idx]


class NormalJumpsPathData(Dataset):
    """as NormalPathData plus per-step jumps (nets.py:178-189)."""

    def __init__(self, paths, payoffs, normals, jumps):
        self.paths = paths[:, :-1]
        self.payoffs = payoffs
        self.normals = normals
        self.jumps = jumps

    def __len__(self):
        return len(self.payoffs)

    def __getitem__(self, idx):
        return (self.paths[idx], self.normals[idx], self.jumps[idx]), self.payoffs[idx]


class AdaptedPathData(Dataset):
    """Jump-adapted trajectories: post-jump states, increments, pre-jump states, times and marks, all without the
    last time index (nets.py:192-207)."""

    def __init__(self, paths, payoffs, normals, left_paths, time_paths, jump_paths, total_steps):
        self.paths = paths[:, :-1]
        self.payoffs = payoffs
        self.normals = normals
        self.left_paths = left_paths[:, :-1]
        self.time_paths = time_paths[:, :-1]
        self.jump_paths = jump_paths[:, :-1]
        self.total_steps = total_steps

    def __len__(self):
        return len(self.payoffs)

    def __getitem__(self, idx):
        return ((self.paths[idx], self.normals[idx], self.left_paths[idx], self.time_paths[idx],
                 self.jump_paths[idx]), self.payoffs[idx])


def get_mlps(problem, num_layers, hidden_size, device):
    """BN-free control-variate nets sized for a Problem: f for the Brownian part, g for the jumps (nets.py:210-219)."""
    d = problem.dim() + 1
    widths = [hidden_size + d] * num_layers
    f = Mlp(d, widths, problem.solver.sde.brown_dim, batch_norm=False, device=device)
    if not problem.solver.has_jumps:
        return f
    return [f, Mlp(d, widths, d - 1, batch_norm=False, device=device)]


def get_opt(models):
    """Adam over one net or over the [f, g] pair (nets.py:222-228)."""
    if isinstance(models, list):
        return optim.Adam(list(models[0].parameters()) + list(models[1].parameters()))
    return optim.Adam(models.parameters())
