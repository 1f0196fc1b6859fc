"""Drop-in for the reference's ``homography`` module: put this directory (and the repository root)
on PYTHONPATH and ``from homography import transformImage`` (reference app.py:9) resolves to the
B200 implementation.  See INTEGRATION.md."""
from ransac_with_homography_b200.homography import *  # noqa: F401,F403
from ransac_with_homography_b200.homography import __all__  # noqa: F401
