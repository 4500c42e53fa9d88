"""audiogan_b200 -- B200-native (sm_100a) implementation of the audiogan GAN training step.

Drop-in `Generator` / `Discriminator` modules and the loop helpers of /root/reference/audiogan.py,
backed by hand-written CUDA kernels behind a C ABI (include/audiogan_b200.h).
"""
from .modules import (Generator, Discriminator, Embedder, binary_cross_entropy_with_logits_per_sample, length_mask,  # noqa: F401
                      calc_dists, fourth_moment, div_roundup, pin_stopper, G_STRUCT, D_STRUCT)
from .train import (FusedRMSprop, check_grad, clip_grad, d_update, g_update, core_step, masked_bce_mean,  # noqa: F401
                    adversarial_movement_d, adversarially_sample_z)
from .graph import GraphedStep  # noqa: F401,E402
from .feed import StepFeed, make_sample, collate, save_checkpoint, load_checkpoint  # noqa: F401,E402
