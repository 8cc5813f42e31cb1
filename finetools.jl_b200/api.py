"""Public names, following FinEtools' exports for the assembly hot path (src/FinEtools.jl:25-673, the subset on the path)."""
from ._lib import FEGPUError, LIB_PATH  # noqa: F401
from .assembly import (AbstractSysmatAssembler, AbstractSysvecAssembler, GPUContext, SysvecAssemblerGPU, makevector,  # noqa: F401
                       SysmatAssemblerSparseDiagGPU, SysmatAssemblerSparseHRZLumpingSymmGPU,
                       SysmatAssemblerFFBlock, SysmatAssemblerSparseGPU,  # noqa: F401
                       SysmatAssemblerSparseSymmGPU, assemble, expectedntriples, makematrix, matrix_blocked_dd,
                       matrix_blocked_df, matrix_blocked_fd, matrix_blocked_ff, setnomatrixresult, startassembly)
from .datacache import DataCache  # noqa: F401
from .femm import (CSys, DeforModelRed3D, FEMMBase, ForceIntensity, bilform_masslike, distribloads, linform_dot,  # noqa: F401
                   bilform_convection, bilform_diffusion, bilform_div_grad, bilform_dot,  # noqa: F401
                   bilform_lin_elastic, innerproduct)  # noqa: F401
from .fesets import (ETYPE, FESET_BY_NAME, FESetH8, FESetH20, FESetH27, FESetQ4, FESetT3, FESetT4, FESetT10)  # noqa: F401
from .fields import (FENodeSet, NodalField, applyebc, gatherdofnums, gathersysvec, gathervalues_asmat, nalldofs, ndofs,  # noqa: F401
                     nents, nfreedofs, numberdofs, setebc)
from .integdomain import IntegDomain, integrationdata, otherdimensionunity  # noqa: F401
from .integrule import GaussRule, TetRule, TriRule  # noqa: F401
from .meshgen import (H8block, H8blockx, H8toH20, H8toH27, H20block, H27block, Q4block, Q4blockx, T3block, T3blockx,  # noqa: F401
                      T4block, T4blockx, T4toT10, T10block, linearspace, meshboundary)
from .parallel import gather_row_blocks, gather_row_blocks_device, owned_ranges_ordered  # noqa: F401
from .partition import pointpartitioning, slab_owner  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
