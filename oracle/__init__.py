"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference hot path.

Nothing in the product package (`medical-sam2_b200/`) may import from here.
Allowed importers: `tests/`, `__graft_entry__.smoke()`, `bench.py` (cpu_baseline /
`--impl reference` legs only).
"""
