"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the VFace denoising hot path plus a
harness that imports the unmodified reference when it is mounted at
/root/reference.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything from here; the product
package (vface_b200/) never does and fails loudly without its CUDA library.

Parity status: the reference ships no tests, golden vectors or fixtures for
this path, so parity is *unpinned by the reference's own
tests*.  The port in oracle/port.py is pinned instead against outputs of the
reference itself, executed in the build container by oracle/make_golden.py and
committed under tests/golden/ (see tests/test_oracle_vs_golden.py), and -- when
/root/reference is present -- directly against the imported reference
(tests/test_oracle_vs_reference.py).
"""
