"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU checker for the CUDA tensor-PLS path: a numpy
restatement of the reference's algorithm (tpls_oracle.py), the restated
tensorly leaves the reference needs in order to be imported at all
(tensorly_standin/), and the script that mints tests/golden/ from the
reference's unmodified source (make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import anything from here.  The product package
(cmtf_pls_b200) never does, and fails loudly when its CUDA library is missing.
"""
