"""CPU oracle for the MoCo-Flow ray-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline.  ``moco_flow_b200`` never imports it.
"""
