"""Host-side helpers with the reference's names (tools/final_util.py of ada-shen/Interpret_quality):
constants :15-19, set_random :113-120, square_distance :134-147, set_model_args :162-204,
set_shapley_batch_size :207-219, set_interaction_batch_size :221-233, load_model :236-262."""
import os
from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..config import CONFIG
from ..models import DGCNN_cls, GCNN_cls, PointConvDensityClsSsg, PointNet2ClsMsg, PointNetCls

NUM_POINTS = 1024
NUM_REGIONS = 32
NUM_SAMPLES_SAVE = 1000
NUM_SAMPLES = 100
K_FOR_DGCNN = 20

MODEL_CLASSES = {
    "pointnet2": PointNet2ClsMsg,
    "pointnet": PointNetCls,
    "dgcnn": DGCNN_cls,
    "gcnn": GCNN_cls,
    "gcnn_adv": GCNN_cls,
    "pointconv": PointConvDensityClsSsg,
}

_CKPT = "checkpoints/exp_MODEL_%s_DATA_%s_POINTNUM_1024_clean/models/model_best.t7"
_CKPT_ADV = "checkpoints/exp_MODEL_gcnn_adv_DATA_%s_POINTNUM_1024_clean_with_all_rot_da/models/model_399.t7"


def mkdir(path):
    os.makedirs(path, exist_ok=True)


class IOStream():
    """Log file + stdout, tools/final_util.py:90-100 of the reference."""

    def __init__(self, path):
        self.f = open(path, 'a')

    def cprint(self, text):
        print(text)
        self.f.write(text + '\n')
        self.f.flush()

    def close(self):
        self.f.close()


def cal_rank(values):
    """0-based ascending rank of every entry (tools/final_util.py:103-106)."""
    return np.argsort(np.argsort(values))


def set_random(seed):
    """Seeds python hashing, numpy's legacy stream and torch exactly like the reference, so that
    replaying the same call sequence afterwards reproduces its permutations / pairs / contexts."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def square_distance(src, dst):
    """Squared distances between 3-d point sets: src (B,N,3), dst (B,M,3) -> (B,N,M)."""
    if src.shape[-1] != 3 or dst.shape[-1] != 3:
        raise ValueError("square_distance: the CUDA path covers 3-d points, the only use on the coalition path")
    return ops.square_distance3(src.contiguous(), dst.contiguous())


def set_model_args(args):
    if args.dataset not in ("modelnet10", "shapenet"):
        raise Exception("Dataset does not exist")
    if args.model not in MODEL_CLASSES:
        raise Exception("Model not implemented")
    if args.model in ("dgcnn", "gcnn", "gcnn_adv"):
        args.k = K_FOR_DGCNN
    if args.model == "pointnet":
        args.feature_transform = True
    args.model_path = (_CKPT_ADV % args.dataset) if args.model == "gcnn_adv" else (_CKPT % (args.model, args.dataset))


def _knob(args, table):
    key = "gcnn" if args.model == "gcnn_adv" else args.model
    if key not in CONFIG[table]:
        raise Exception("Not implemented")
    return CONFIG[table][key]


def set_shapley_batch_size(args):
    args.shapley_batch_size = _knob(args, "shapley_batch_size")


def set_interaction_batch_size(args):
    args.interaction_batch_size = _knob(args, "interaction_batch_size")


def strip_module_prefix(state_dict):
    """Checkpoints saved from nn.DataParallel carry a 'module.' prefix."""
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[len("module."):] if "module." in k else k] = v
    return out


def build_model(args, state_dict=None):
    """Model of args.model on args.device in eval mode, optionally from an in-memory state dict."""
    if args.model not in MODEL_CLASSES:
        raise Exception("Not implemented")
    model = MODEL_CLASSES[args.model](args).to(args.device)
    if state_dict is not None:
        sd = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v)))
              for k, v in strip_module_prefix(state_dict).items()}
        model.load_state_dict(sd)
    return model.eval()


def load_model(args):
    state_dict = torch.load(args.model_path, map_location=args.device)
    return build_model(args, state_dict)
