"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the B200 coupling engine: the unmodified reference particle path
(`oracle/_ref/libfoamyade_ref.so`, wrapped by `oracle.ref`) and this repo's CPU
restatement of the particle + FV/PISO halves (`oracle/_build/liboracle.so`,
wrapped by `oracle.port`).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product
(`yade-openfoam-coupling_b200/`) never does.
"""
