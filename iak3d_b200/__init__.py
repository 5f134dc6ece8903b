"""iak3d_b200 -- B200 (sm_100a) implementation of iak3d's data-parallel hot path behind the reference's interface.

The compute lives in lib/libiak3d.so (CUDA, C ABI: include/iak3d.h); this package is the host-side mirror of the
reference's R functions for that path.  Importing the package does not load the library; the first call does, and
fails loudly when the library or an sm_100-class GPU is missing (there is no CPU path).
"""
from ._lib import EXPORTS, LIB_PATH, Context, Iak3dError, default_context, load  # noqa: F401
from .api import (calcbetavXbeta, lnN_design, makeXnsClampedUG0, nrUpdatesIAK3DlnN_gram, setmuIAK3D, nrUpdatesIAK3D, nr_iterate, predictIAK3D_CL, profilePredictIAK3D, predMatsIAK3D_CL_ByBlock, predMatsIAK3D_CLEV_ByBlock, getBlocksFromLocations,  # noqa: F401
                  BADNLL, KeptFit, SetupMats, SetupMats2, cl_plan_units, cl_reduce_and_tail, gradHessIAK3D, gradnllIAK3D, iaCovMatern2, lndetANDinvCb, logitab,  # noqa: F401
                  nll_terms_batch, nllIAK3D, nllIAK3D_CL, gradnllIAK3D_CL, predictIAK3D, readParsIAK3D, reml_tail, reml_tail_batch, setCIAK3D, setCIAK3D2,
                  setupIAK3D, setupIAK3D2, setupIAK3D_CL, spatialCovIAK3D, tauTfm)
from .devmat import DeviceMat, cov_mat  # noqa: F401
from .design import DesignModel, makeXvX  # noqa: F401
from .blocks import (DevicePoints, balanceNeighbours, balanceNeighbours2, calcnPerCluster, setVoronoiBlocks,  # noqa: F401
                     setVoronoiBlocksWrap)
from . import parallel, synth  # noqa: F401
