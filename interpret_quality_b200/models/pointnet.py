"""PointNetCls with the reference's constructor, checkpoint keys and 3-tuple forward
(models/pointnet.py:91-115 of ada-shen/Interpret_quality); runs in csrc/pointnet_model.cu."""
from ._base import IQModule


class PointNetCls(IQModule):
    KIND = "pointnet"

    def __init__(self, args):
        if not getattr(args, "feature_transform", True):
            raise NotImplementedError("only feature_transform=True is supported (set_model_args always sets it, "
                                      "tools/final_util.py:171,193)")
        super().__init__(args)
        self.feature_transform = True

    def forward(self, x):
        """-> (logits (B,C), trans_feat (B,64,64), crt_points (B,1024) int64), models/pointnet.py:115."""
        return self._run(x, point_major=False, want_aux=True)
