"""Import stand-in for matplotlib (utils.py:41): the trainer's figure calls become no-ops."""
