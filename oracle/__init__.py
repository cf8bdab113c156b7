"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's algorithm for the simulation hot path (see
``oracle/pouch_oracle.py`` and ``oracle/csrc/oracle.c``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker or the reported CPU baseline.  The product
(``calcium_b200/``) never imports it and fails loudly without its CUDA library.
"""
