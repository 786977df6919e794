"""RFI flagging: device operations (``device``) and the host-call interfaces (``host``)."""

#: Ratio of the standard deviation to the median absolute deviation of a normal
#: distribution, as the reference defines it (``rfi/__init__.py:31``).
MAD_NORMAL = 1.4826
