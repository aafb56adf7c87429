"""Built-in models and the model-callable conventions of mc3_b200.

The reference's `func` is any Python callable `func(params, *indparams,
**indparams_dict) -> 1-D array` evaluated once per chain-step
(mc3/chain.py:316-319).  mc3_b200 accepts three kinds of `func`:

* a BuiltinModel (below): evaluated inside the fused CUDA model+chi-squared
  kernel for all chains at once -- the fast path;
* a TorchModel: a user torch callable evaluated ONCE per generation over the
  whole population, `fn(P[nchains, npars], *indparams) -> model[nchains, N]`
  on the device;
* any other callable: the reference's own convention, evaluated per chain on
  the host (compatibility path; the chi-squared still runs on the GPU).

BuiltinModel objects are also plain callables with the reference signature
`model(params, x)`; the evaluation runs on the GPU (mc3b_model_eval).
"""
import numpy as np

from . import _lib

POLYNOMIAL, SINUSOID, GAUSSIAN, BOX, SINUSOID_GRID = 0, 1, 2, 3, 4


class BuiltinModel:
    """A model implemented in mc3_b200/csrc/models.cuh.

    polynomial  y = sum_k p[k] x^k            (1..8 coefficients)
    sinusoid    y = p0 sin(2 pi x/p1 + p2) + p3 + p4 x
    gaussian    y = p0 exp(-0.5 ((x-p1)/p2)^2) + p3
    box         y = p3 - p0 [|x-p1| < p2/2]
    """

    def __init__(self, name, model_id, nparams):
        self.name, self.model_id, self.nparams = name, model_id, nparams

    def nmodel(self, nfunc_params):
        """Number of leading parameters the model consumes."""
        if self.nparams is None:
            if not 1 <= nfunc_params <= 8:
                raise ValueError(
                    f'{self.name} takes 1 to 8 parameters, got {nfunc_params}')
            return nfunc_params
        if nfunc_params != self.nparams:
            raise ValueError(
                f'{self.name} takes {self.nparams} parameters, got {nfunc_params}')
        return self.nparams

    def __call__(self, params, x):
        import torch
        params = np.atleast_2d(np.asarray(params, dtype=np.double))
        x = np.ascontiguousarray(x, dtype=np.double)
        nmodel = self.nmodel(params.shape[1])
        dev = torch.device('cuda')
        dp = torch.from_numpy(np.ascontiguousarray(params)).to(dev)
        dx = torch.from_numpy(x).to(dev)
        out = torch.empty((params.shape[0], x.size), dtype=torch.float64, device=dev)
        _lib.call('mc3b_model_eval', self.model_id, dp.data_ptr(), params.shape[1],
                  params.shape[0], nmodel, dx.data_ptr(), x.size, out.data_ptr(),
                  _lib.stream_ptr())
        res = out.cpu().numpy()
        return res[0] if res.shape[0] == 1 else res

    def __repr__(self):
        return f'<mc3_b200 built-in model {self.name}>'


class TorchModel:
    """Marks `fn` as batched over chains: fn(P[nchains, npars], *indparams,
    **indparams_dict) -> torch tensor [nchains, N] (float64, on P's device)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, params, *args, **kwargs):
        import torch
        p = torch.as_tensor(np.atleast_2d(np.asarray(params, dtype=np.double)),
                            device='cuda')
        args = [torch.as_tensor(a, device='cuda') if isinstance(a, np.ndarray) else a
                for a in args]
        out = self.fn(p, *args, **kwargs).detach().cpu().numpy()
        return out[0] if np.ndim(params) == 1 else out


polynomial = BuiltinModel('polynomial', POLYNOMIAL, None)
sinusoid = BuiltinModel('sinusoid', SINUSOID, 5)
gaussian = BuiltinModel('gaussian', GAUSSIAN, 4)
box = BuiltinModel('box', BOX, 4)
BUILTIN = {'polynomial': polynomial, 'sinusoid': sinusoid,
           'gaussian': gaussian, 'box': box}
