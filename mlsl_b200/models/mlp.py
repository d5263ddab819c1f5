import torch


class MLP(torch.nn.Module):
    def __init__(self, d_in=1024, d_hidden=4096, d_out=1024, layers=4):
        super().__init__()
        dims = [d_in] + [d_hidden] * (layers - 1) + [d_out]
        mods = []
        for i in range(layers):
            mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
            if i + 1 < layers:
                mods.append(torch.nn.GELU())
        self.net = torch.nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)
