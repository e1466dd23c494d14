"""TEST INFRASTRUCTURE — not part of the product.

`oracle/` is the CPU checker for the B200 hot path: a restatement of the reference's algorithm in
plain torch-CPU fp32 / numpy (`med3d_oracle.py`, `pipeline_oracle.py`), the seeded synthetic
inputs and weights both sides load (`synthetic.py`), and — usable only where /root/reference is
mounted — a shim that imports the unmodified reference to pin the restatement and to generate the
golden vectors under tests/golden/ (`ref_shim.py`, `make_golden.py`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under bodyct-dram-emph-subtype_b200/ does.
"""
