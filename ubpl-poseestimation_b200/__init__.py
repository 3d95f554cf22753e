"""ubpl_b200 -- B200-native (sm_100a) pseudo-label hot path of Qi2019KB/UBPL-PoseEstimation."""
__version__ = "0.1.0"
