"""CPU oracles for the RFI flagging hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``katsdpsigproc_b200`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker or as the reported CPU
baseline -- never as the thing shipped.

Two tiers ("Oracle"):

``oracle.host_numpy``
    numpy/pandas restatement of the reference's ``katsdpsigproc.rfi.host``
    classes (float64, same library calls, same operation order).  Pinned
    against the reference's own known-answer tests and against outputs of the
    real reference generated in the build container
    (``tests/golden/make_golden.py``).

``oracle.contract``
    plain-C restatement (``contract.c``, built with gcc) of the *float32
    device contract*: the exact arithmetic the CUDA kernels promise (rules
    R1-R10 of SURVEY.md section 8(c')).  The CUDA path must match it bit for
    bit; it must match ``host_numpy`` as the rules say (selections exact,
    noise within 1 ulp, flags exact outside a counted 1e-6 band).
"""
