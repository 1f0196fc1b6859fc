"""Drop-in for the reference's ``ransac`` module (``from ransac import stitching``, reference
app.py:10; ``from ransac import HomoModel, RANSAC, calcHomographyLinear, calcHomography``,
regression.ipynb:561).  See INTEGRATION.md."""
from ransac_with_homography_b200.ransac import (DEBUG, LVL, Model, HomoModel, RANSAC, example0, stitching,  # noqa: F401
                                                calcHomographyLinear, calcHomography, stitchPanorama, cylindericlMap)
