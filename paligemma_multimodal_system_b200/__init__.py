"""B200-native PaliGemma inference (sm_100a kernels behind the reference's Python module API)."""
