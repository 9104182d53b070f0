"""PointNet2ClsMsg with the reference's constructor and checkpoint keys
(models/pointnet2.py:244-276 of ada-shen/Interpret_quality)."""
from ._base import IQModule


class PointNet2ClsMsg(IQModule):
    KIND = "pointnet2"
