"""B200-native implementation of katsdpsigproc's RFI-flagging hot path.

Mirrors the reference's module layout for the path (``accel``, ``cuda``,
``transpose``, ``percentile``, ``maskedsum``, ``rfi.device``); all device work
is done by hand-written sm_100a kernels in ``_lib/libksp_b200.so`` reached
through the C ABI of ``include/ksp_b200.h``.  There is no CPU fallback.
"""

__version__ = "0.1.0"
