from .dgcnn import DGCNN_cls, GCNN_cls  # noqa: F401
from .pointconv import PointConvDensityClsSsg  # noqa: F401
from .pointnet import PointNetCls  # noqa: F401
from .pointnet2 import PointNet2ClsMsg  # noqa: F401
