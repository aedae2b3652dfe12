"""CPU parity oracle for the rasterizer hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__ (build / smoke checker) and bench.py's CPU
baseline legs may import this package; the product package never does.
See oracle/splat_oracle.c (C restatement of the reference CUDA kernels) and
oracle/torch_oracle.py (autograd restatement of the reference's torch path).
"""
