"""interpret_quality_b200 -- B200-native coalition evaluation for ada-shen/Interpret_quality.

Host-side mirror of the reference's Python interface for ONE hot path (FPS regions -> coalition
masking -> masked forward -> Shapley / interaction reduction), backed by hand-written sm_100a CUDA
kernels behind the C ABI of include/iq_b200.h.  Module names follow the reference:

    interpret_quality_b200.models.{pointnet,pointnet2,pointconv,dgcnn}
    interpret_quality_b200.tools.{final_util,final_common}
    interpret_quality_b200.{final_save_fps,final_shapley_value,
                            final_point_binary_interaction_logits,final_cal_interactions,config}
"""
__version__ = "0.1.0"
