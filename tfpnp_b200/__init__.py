"""tfpnp_b200 -- B200-native PnP-ADMM inner solver behind the TFPnP solver API.

Host-side mirror of the reference interface for ONE hot path:
``PnPSolver.forward(inputs, parameters)`` for CS-MRI / phase retrieval / sparse-view CT /
single-photon imaging, with the UNet prox_sigma denoiser, implemented as hand-written
sm_100a CUDA in ``libtfpnp_b200.so`` (C ABI: include/tfpnp_b200.h).  No fallback paths.
"""
from ._lib import build, lib, LIB_PATH, release_cached_scratch  # noqa: F401
from .denoiser import UNetDenoiser2D, IRCNNDenoiser2D, create_denoiser, random_unet_state_dict  # noqa: F401
from .solver import (PnPSolver, ADMMSolver, IADMMSolver, ADMMSolver_CSMRI, IADMMSolver_PR,  # noqa: F401
                     HQSSolver_CSMRI, PGSolver_CSMRI, APGSolver_CSMRI, REDADMMSolver_CSMRI, PGSolver_CT,
                     IADMMSolver_CT, ADMMSolver_SPI, RadonGenerator, create_solver_csmri,
                     create_solver_pr, create_solver_ct, create_solver_spi)
from .ops import (radon_forward, radon_backward, torch_psnr, conv3x3_lrelu_nhwc, fft2, ifft2, complex_mul,  # noqa: F401
                  cdp_forward, cdp_backward)
from .measure import csmri_measure, pr_measure, spi_measure, ct_measure, radial_mask  # noqa: F401
from .dist import shard_batch, shard_bounds, all_gather_psnr, all_gather_batch, NativeComm  # noqa: F401
from .env import Batch, Env, DifferentiableEnv, PnPEnv, CSMRIEnv, PREnv, CTEnv, SPIEnv  # noqa: F401
