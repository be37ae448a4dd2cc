"""Plan caching shared by the backbone modules and CA_PF."""
import os

import torch

from ... import lib, program


def default_precision():
    return os.environ.get("CAPF_PRECISION", "fp32")


def default_use_tc():
    """tcgen05/TMA kernels for every 16-bit GEMM-shaped op (CAPF_TCGEN05=0 forces the CUDA-core kernels)."""
    return os.environ.get("CAPF_TCGEN05", "1") != "0"


def state_version(module: torch.nn.Module):
    """Change detector for the packed weights of a plan: a fingerprint of every parameter and buffer -- identity of the
    tensor object, address of its storage and its autograd version counter.  It changes when a tensor is replaced
    (``load_state_dict(assign=True)``, ``.to()``), re-allocated, or written in place through the tensor itself
    (``load_state_dict``, ``p.copy_()``, optimiser steps).  Writes that bypass the version counter -- ``p.data.add_(1)``,
    ``p.data.copy_(w)`` -- are invisible to any cheap check: call ``invalidate_weights()`` on the model after such edits."""
    return hash(tuple((id(t), t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers())))


class PlanCacheMixin:
    """Device plans hold ctypes handles, CUDA streams and events: they are a cache, not state.  They are dropped when the
    parameters move (``_apply``), never copied or pickled (``copy.deepcopy(model)`` / ``torch.save(model)`` work after a
    forward), and ``invalidate_weights()`` forces the next forward to repack every weight blob."""

    def invalidate_weights(self):
        """Force a repack of the packed device weights on the next forward.  Needed only after edits that bypass the
        tensors' version counters (``p.data.<op>_()``, raw pointer writes); every other change is detected."""
        for ent in self.__dict__.get("_plans", {}).values():
            ent[1] = None
        for m in self.children():
            if isinstance(m, PlanCacheMixin):
                m.invalidate_weights()

    def _apply(self, fn, *a, **k):
        self.__dict__["_plans"] = {}            # .to()/.cuda()/.half() move the parameters: drop device plans
        return super()._apply(fn, *a, **k)

    def __getstate__(self):
        st = dict(super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__)
        st["_plans"] = {}
        return st

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_plans" else copy.deepcopy(v, memo)
        return new


class BackboneRuntimeMixin(PlanCacheMixin):
    """Lets a backbone be called on its own: NCHW in -> list of 4 NCHW maps (reference signature)."""

    def forward(self, x):
        if not x.is_cuda:
            raise lib.CapfError("the backbone runs on a B200 through libcapf_b200; got a CPU tensor (no CPU path)")
        B, C, H, W = x.shape
        prec = getattr(self, "precision", None) or default_precision()
        key = (B, H, W, prec, x.device.index)
        cache = self.__dict__.setdefault("_plans", {})
        state = {"backbone." + k: v for k, v in self.state_dict().items()}
        ver = state_version(self)
        ent = cache.get(key)
        if ent is None:
            shapes = {k: tuple(v.shape) for k, v in state.items()}
            bb = "cpn" if self.kind == "cpn" else "hrnet_32"
            prog = program.build_forward_program(bb, getattr(self, "cfg", None), None, shapes, B, H, W, prec,
                                                 use_tc=default_use_tc(), backbone_only=True)
            ent = [program.Plan(prog, state, x.device), ver]
            cache[key] = ent
        elif ent[1] != ver:
            ent[0].repack(state)
            ent[1] = ver
        plan = ent[0]
        plan.tensor(plan.prog.inputs["images"]).copy_(x.permute(0, 2, 3, 1))
        plan.run()
        return [plan.tensor(m).permute(0, 3, 1, 2).float() for m in plan.prog.feature_maps]
