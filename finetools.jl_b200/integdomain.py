"""IntegDomain: finite elements + quadrature rule (reference: src/IntegDomainModule.jl:43-61, 150-152, 610-648)."""
import numpy as np


def otherdimensionunity(loc=None, conn=None, N=None):
    return 1.0


class IntegDomain:
    def __init__(self, fes, integration_rule, otherdimension=None, axisymmetric=False):
        self.fes = fes
        self.integration_rule = integration_rule
        # a constant "other dimension" (IntegDomainModule.jl:73-82) is the only kind that can cross to the GPU
        if otherdimension is None:
            self.otherdimension = 1.0
        elif callable(otherdimension):
            if otherdimension is otherdimensionunity:
                self.otherdimension = 1.0
            else:
                raise ValueError("only a constant other-dimension is GPU-eligible (no callbacks cross the C ABI)")
        else:
            self.otherdimension = float(otherdimension)
        self.axisymmetric = bool(axisymmetric)


def integrationdata(integdomain, integration_rule=None):
    """npts, Ns[j] (nne x 1), gradNparams[j] (nne x mdim), w (npts x 1), pc (npts x mdim)."""
    rule = integration_rule if integration_rule is not None else integdomain.integration_rule
    pc = np.asarray(rule.param_coords, dtype=np.float64)
    w = np.asarray(rule.weights, dtype=np.float64)
    npts = rule.npts
    Ns = [integdomain.fes.bfun(pc[j, :]) for j in range(npts)]
    gradNparams = [integdomain.fes.bfundpar(pc[j, :]) for j in range(npts)]
    return npts, Ns, gradNparams, w, pc
