"""kd_6d_pose_adlp_b200 -- B200-native optimal-transport knowledge-distillation path.

A from-scratch implementation of ONE hot path of GUOShuxuan/kd-6d-pose-adlp: the OT distillation loss
(``losses/kd_loss.py`` -> ``losses/loss_libs.py`` -> ``geomloss.SamplesLoss``) and the teacher-side cell
selection (``postprocess/postprocess_kd.py``), as hand-written sm_100a CUDA kernels behind a C ABI
(``include/kdot.h``, ``libkdot.so``) with a thin Python host side that mirrors the reference's interface:

=====================================  ==========================================================
reference                              this package
=====================================  ==========================================================
``geomloss.SamplesLoss``               :class:`kd_6d_pose_adlp_b200.samples_loss.SamplesLoss`
``losses.loss_libs.kd_loss_2d``        :func:`kd_6d_pose_adlp_b200.losses.loss_libs.kd_loss_2d`
``losses.kd_loss.KDPoseLoss``          :class:`kd_6d_pose_adlp_b200.losses.kd_loss.KDPoseLoss`
``postprocess.postprocess_kd``         :class:`kd_6d_pose_adlp_b200.postprocess.postprocess_kd.PostProcessorKD`
=====================================  ==========================================================

There is no CPU fallback and no Triton / torch.compile path: without ``libkdot.so`` and a CUDA device every
operator raises.
"""
from .ops import OTConfig, OTLossFunction, ot_loss_batched  # noqa: F401
from .samples_loss import SamplesLoss  # noqa: F401

__version__ = "0.1.0"
