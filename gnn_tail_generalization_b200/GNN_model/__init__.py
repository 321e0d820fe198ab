"""Drop-in replacement for the reference's ``GNN_model`` package (TeacherGNN path only).

Put ``<repo>/gnn_tail_generalization_b200`` ahead of the reference checkout on ``sys.path`` /
``PYTHONPATH`` and the reference's unchanged ``trainer_node_classification.py`` resolves
``from GNN_model.GNN_normalizations import TeacherGNN`` to the B200-native modules in this directory
(same class names, constructor signatures, attribute chain and ``state_dict`` keys).
It can also be imported as ``gnn_tail_generalization_b200.GNN_model``.
"""
