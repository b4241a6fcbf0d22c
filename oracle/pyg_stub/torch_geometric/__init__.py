"""Minimal stand-in for torch_geometric 2.1.0 (requirements.txt:49 of the reference).

TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own, unmodified
``Model/LightGCN.py`` / ``Model/LATTICE.py`` can be imported in the authoring container
(where torch_geometric is not installed) to pin ``oracle/ref_ops.py`` and to generate the
fixtures under ``tests/golden``.  It restates the published gather -> message ->
scatter-add semantics of ``MessagePassing(aggr='add')`` and ``utils.degree``.
"""
