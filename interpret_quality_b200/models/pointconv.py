"""PointConvDensityClsSsg with the reference's constructor and checkpoint keys
(models/pointconv.py:394-424 of ada-shen/Interpret_quality)."""
from ._base import IQModule


class PointConvDensityClsSsg(IQModule):
    KIND = "pointconv"
