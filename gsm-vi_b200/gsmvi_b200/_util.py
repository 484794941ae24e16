"""Host-side helpers shared by the GSM / BaM drivers: device buffers with TMA-friendly padding, key handling."""
import numpy as np
import torch

from ._lib import GsmviError

PAD = 32  # leading dimensions are multiples of 32 floats (128 B rows: TMA needs 16 B, full sectors are nicer)


def ld_of(D):
    return (D + PAD - 1) // PAD * PAD


def device():
    if not torch.cuda.is_available():
        raise GsmviError("gsmvi_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def new_mat(rows, D, dev):
    """rows x D fp32 matrix stored with a padded leading dimension; returns (storage, view[:, :D])."""
    buf = torch.zeros(rows, ld_of(D), dtype=torch.float32, device=dev)
    return buf, buf[:, :D]


def new_vec(D, dev):
    return torch.zeros(ld_of(D), dtype=torch.float32, device=dev)


def to_dev(a, dev):
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=dev, dtype=torch.float32)
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(dev)


def key_to_seed(key):
    """Reference keys are a JAX PRNGKey (uint32[2], gsmvi/gsm.py:117) or an int (gsmvi/gsm_numpy.py:105).  Any of
    int / 2-word array / torch.Generator is mapped to a 64-bit Philox seed.  Bit-exact parity with the reference RNG
    stream (threefry + MT19937 + SVD transform) is not a goal."""
    if isinstance(key, torch.Generator):
        return int(key.initial_seed()) & (2**64 - 1)
    if isinstance(key, (int, np.integer)):
        return int(key) & (2**64 - 1)
    arr = np.asarray(key.cpu() if isinstance(key, torch.Tensor) else key).astype(np.uint64).ravel()
    if arr.size == 1:
        return int(arr[0])
    return (int(arr[0]) << 32 | int(arr[1])) & (2**64 - 1)
