"""Import stand-in for torch_geometric (see shims/README.md): only what the TeacherGNN path of the reference's
unchanged trainer touches."""
from . import data, datasets, transforms, utils, typing  # noqa: F401

__version__ = '0.0-shim'
