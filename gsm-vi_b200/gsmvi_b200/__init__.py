"""gsmvi_b200: B200-native GSM / BaM variational inference (hot path of modichirag/GSM-VI).

    from gsmvi_b200.gsm import GSM, gsm_update
    from gsmvi_b200.targets import DenseGaussianTarget

Everything numerical runs in libgsmvi_b200.so (hand-written sm_100a CUDA, see include/gsmvi_b200.h); importing the
package does not load it, calling any operator does and fails loudly if it is missing."""
__version__ = "0.1.0"
