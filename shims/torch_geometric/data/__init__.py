from . import data  # noqa: F401  (utils.py:798 names torch_geometric.data.data.Data)
from .data import Data, DataLoader  # noqa: F401
