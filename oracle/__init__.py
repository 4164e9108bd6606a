"""CPU oracle for the FIGAROH dynamic-identification path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``figaroh-plus_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do.  See oracle/README.md.
"""
