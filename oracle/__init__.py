"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (pn2_oracle.c) of the reference's pointnet_lib kernels, the
build recipe for the reference's own kernels (oracle/_ref, Makefile) and thin
loaders for both.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / reference legs may import this package; hotrack_b200/ never does.
"""
