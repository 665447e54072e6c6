"""megba_b200 -- Blackwell-native LM / Schur-PCG hot path behind MegBA's API.

Python is only the test / benchmark harness: the product is the C-ABI shared library
built from megba_b200/csrc (see include/megba_b200.h) and the C++ facade in
include/megba/.  Importing this package never falls back to a CPU implementation.
"""
