"""DGCNN_cls / GCNN_cls with the reference's constructor, checkpoint keys and forward contract
(models/dgcnn.py:51-120 and :123-194 of ada-shen/Interpret_quality); the forward pass runs in
csrc/edgeconv_model.cu."""
from ._base import IQModule


class DGCNN_cls(IQModule):
    """EdgeConv classifier with the kNN graph rebuilt in feature space before every layer; args.k neighbours."""
    KIND = "dgcnn"

    def __init__(self, args):
        super().__init__(args)
        self.k = int(args.k)


class GCNN_cls(IQModule):
    """Same network with the graph fixed on the input coordinates (models/dgcnn.py:161)."""
    KIND = "gcnn"

    def __init__(self, args):
        super().__init__(args)
        self.k = int(args.k)
