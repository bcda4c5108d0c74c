"""CPU oracle for the EPC-Net embedding/retrieval path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``epc-net_b200/`` imports this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs.
"""
