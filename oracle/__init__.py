"""ORACLE -- test infrastructure, not product code.

A CPU restatement of the reference's coalition-evaluation path
(ada-shen/Interpret_quality), used as the checker by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (interpret_quality_b200/) never imports it.

* geom_oracle.c  -- plain C, integer-valued geometry (FPS, nearest centre, ball
                    query, coalition masks), every rounding spelled out.
* coalition.py   -- numpy restatement of reward, Shapley and interaction sums.
* nets.py        -- torch fp32 (CPU) functional restatement of the five
                    classifiers, driven by a checkpoint-format state dict.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself,
imported unmodified from /root/reference in the build container by
tests/golden/make_golden.py; the resulting fixtures live in tests/golden/*.npz
and tests/test_oracle_golden.py holds the oracle to them.
"""
