"""Exception types of the reference's lower seam that callers may catch."""


class DGLError(Exception):
    """Raised where the reference raised ``dgl.base.DGLError`` (GCN.py:187-197, 215-219)."""
