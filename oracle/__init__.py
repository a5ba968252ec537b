"""TEST INFRASTRUCTURE ONLY. CPU restatements of the reference algorithms + a ctypes wrapper of
the unmodified reference CUDA core (oracle/_ref). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package."""
