"""CPU oracle for the AdaptSegNet output-space-adaptation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``adaptsegnet_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the product path.

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's own
modules imported from ``/root/reference`` in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
