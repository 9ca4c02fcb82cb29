"""CPU oracle for the PREGO MiniROAD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``prego_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker or
as the CPU baseline -- never as the product path.
"""
