"""Library version, ``major.minor.patch`` of the reference this engine is a drop-in for
(ref ``nls/version.py:5-12``, ``nls/nls.f90:29-37``)."""

MAJOR, MINOR, PATCH = 0, 2, 0


def version():
    return MAJOR, MINOR, PATCH
