"""absl flags with the reference's names and defaults for the hot path.

Reference definitions: frame_level_models.py:35-53,2197-2216; video_level_models.py:26-45;
train.py:44-112.  `tf.flags` in TF 1.x is absl.flags, so the same parsing rules apply.
"""
from absl import flags

FLAGS = flags.FLAGS


def _define(fn, name, default, help_):
    if name not in FLAGS:
        fn(name, default, help_)


# frame_level_models.py:35-42
_define(flags.DEFINE_integer, "iterations", 30, "Number of frames per batch for DBoF.")
_define(flags.DEFINE_bool, "sample_random_frames", True,
        "If true samples random frames (for frame level models). Unused by NetVladV1/V2; read by WillowModelReg.")
# frame_level_models.py:2197-2216
_define(flags.DEFINE_bool, "netvlad_add_batch_norm", True, "Adds batch normalization to the DBoF model.")
_define(flags.DEFINE_integer, "netvlad_cluster_size", 256, "Number of units in the NetVLAD cluster layer.")
_define(flags.DEFINE_integer, "netvlad_hidden_size", 1024, "Number of units in the NetVLAD hidden layer.")
_define(flags.DEFINE_bool, "netvlad_relu", False, "add ReLU to hidden layer")
_define(flags.DEFINE_bool, "gating", True, "Gating for NetVLAD")
_define(flags.DEFINE_bool, "gating_remove_diag", False, "Remove diag for self gating")
_define(flags.DEFINE_float, "audio_det_reg", 1e-4,
        "The coefficient that determines the strength of the determinant regularization penalty "
        "(of the VLAD cluster centres, for audio features).")
_define(flags.DEFINE_float, "rgb_det_reg", 1e-4,
        "The coefficient that determines the strength of the determinant regularization penalty "
        "(of the VLAD cluster centres, for rgb features).")
# video_level_models.py:26-45
_define(flags.DEFINE_integer, "moe_num_mixtures", 2,
        "The number of mixtures (excluding the dummy 'expert') used for MoeModel.")
_define(flags.DEFINE_float, "moe_l2", 1e-8, "L2 penalty for MoeModel.")
_define(flags.DEFINE_integer, "moe_low_rank_gating", -1, "Low rank gating for MoeModel.")
_define(flags.DEFINE_bool, "moe_prob_gating", False, "Prob gating for MoeModel.")
_define(flags.DEFINE_string, "moe_prob_gating_input", "prob", "input Prob gating for MoeModel.")
# train.py:44-112 (trainer flags on the hot path)
_define(flags.DEFINE_string, "model", "NetVladV1", "Which architecture to use for the model.")
_define(flags.DEFINE_integer, "num_gpu", 1, "The maximum number of GPU devices to use for training.")
_define(flags.DEFINE_integer, "batch_size", 1024, "How many examples to process per batch (per tower).")
_define(flags.DEFINE_string, "label_loss", "CrossEntropyLoss", "Which loss function to use.")
_define(flags.DEFINE_float, "regularization_penalty", 1.0, "Weight of the regularization loss.")
_define(flags.DEFINE_float, "base_learning_rate", 0.01, "Which learning rate to start with.")
_define(flags.DEFINE_float, "learning_rate_decay", 0.95, "Learning rate decay factor.")
_define(flags.DEFINE_float, "learning_rate_decay_examples", 4000000, "Decay every this many examples.")
_define(flags.DEFINE_float, "clip_gradient_norm", 1.0, "Norm to clip gradients to.")
_define(flags.DEFINE_string, "optimizer", "AdamOptimizer", "What optimizer class to use.")
_define(flags.DEFINE_bool, "frame_features", True, "Frame-level features.")
_define(flags.DEFINE_string, "feature_names", "rgb,audio", "Names of the features.")
_define(flags.DEFINE_string, "feature_sizes", "1024,128", "Lengths of the feature vectors.")
_define(flags.DEFINE_integer, "max_steps", None, "The maximum number of iterations of the training loop.")


def ensure_parsed():
    """Flags are read at create_model time (like the reference); parse defaults if nobody called app.run."""
    if not FLAGS.is_parsed():
        FLAGS.mark_as_parsed()
