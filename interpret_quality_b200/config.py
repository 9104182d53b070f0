# Batch knobs with the reference's names and defaults (config.py:1-16 of ada-shen/Interpret_quality).
# They bound how many permutations / contexts the caller hands over per step; the engine may fuse or
# split them internally, results do not depend on them.
CONFIG = {
    "shapley_batch_size": {"pointnet2": 5, "pointnet": 50, "dgcnn": 5, "gcnn": 10, "pointconv": 20},
    "interaction_batch_size": {"pointnet2": 25, "pointnet": 100, "dgcnn": 25, "gcnn": 50, "pointconv": 100},
}
