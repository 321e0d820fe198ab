"""Import stand-in for ogb (utils.py:29); ogbn-arxiv itself is not on disk."""
