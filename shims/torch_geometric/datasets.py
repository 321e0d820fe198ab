"""Dataset classes of the shim: the real Planetoid / WebKB / ... files are not on disk and cannot be downloaded, so
every class returns a seeded synthetic graph with the named dataset's shape (N, F, C from base_options.py:187-300):
bag-of-words-like non-negative features, labels correlated with them, a power-law undirected graph without self
loops, Planetoid-style masks."""
import torch

from .data import Data

_SHAPES = {  # name: (N, F, C, undirected edges)
    'cora': (2708, 1433, 7, 5278), 'citeseer': (3327, 3703, 6, 4552), 'pubmed': (19717, 500, 3, 44324),
    'texas': (183, 1703, 5, 295), 'wisconsin': (251, 1703, 5, 466), 'cornell': (183, 1703, 5, 280),
    'actor': (7600, 932, 5, 26752), 'chameleon': (2277, 128, 6, 31421), 'squirrel': (5201, 128, 5, 198493),
}


def synthetic(name, seed=0):
    n, f, c, und = _SHAPES[name.lower()]
    g = torch.Generator().manual_seed(1234 + seed)
    y = torch.randint(0, c, (n,), generator=g)
    proto = (torch.rand(c, f, generator=g) < 0.02).float()
    x = ((torch.rand(n, f, generator=g) < 0.01).float() + proto[y] * (torch.rand(n, f, generator=g) < 0.5).float()).clamp(max=1)
    x[x.sum(1) == 0, 0] = 1.0
    w = torch.arange(1, n + 1, dtype=torch.float64).pow(-0.6)
    perm = torch.randperm(n, generator=g)
    a = perm[torch.multinomial(w, 2 * und, replacement=True, generator=g)]
    b = perm[torch.multinomial(w, 2 * und, replacement=True, generator=g)]
    keep = a != b
    lo, hi = torch.minimum(a[keep], b[keep]), torch.maximum(a[keep], b[keep])
    key = torch.unique(lo * n + hi)[:und]
    lo, hi = key // n, key % n
    # every node gets at least one neighbour (Planetoid graphs have no isolated nodes to speak of)
    deg = torch.bincount(torch.cat([lo, hi]), minlength=n)
    lone = (deg == 0).nonzero().flatten()
    lo, hi = torch.cat([lo, lone]), torch.cat([hi, (lone + 1) % n])
    edge_index = torch.stack([torch.cat([lo, hi]), torch.cat([hi, lo])])
    idx = torch.randperm(n, generator=g)
    train_mask, val_mask, test_mask = (torch.zeros(n, dtype=torch.bool) for _ in range(3))
    train_mask[idx[: 20 * c]] = True
    val_mask[idx[20 * c: 20 * c + 500]] = True
    test_mask[idx[-1000:]] = True
    return Data(x=x, y=y, edge_index=edge_index, train_mask=train_mask, val_mask=val_mask, test_mask=test_mask)


class _Synthetic:
    def __init__(self, root, name=None, split='public', transform=None, **kw):
        self.name = name if name is not None else type(self).__name__
        self.transform = transform

    def __getitem__(self, i):
        d = synthetic(self.name)
        return self.transform(d) if self.transform is not None else d

    def __len__(self):
        return 1


class Planetoid(_Synthetic):
    pass


class WebKB(_Synthetic):
    pass


class WikipediaNetwork(_Synthetic):
    pass


class Actor(_Synthetic):
    def __init__(self, root, transform=None, **kw):
        super().__init__(root, 'actor', transform=transform)


class Coauthor(_Synthetic):
    pass


class Amazon(_Synthetic):
    pass
