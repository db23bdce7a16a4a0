"""Flat parameter / gradient storage for layers with many small dense float32 parameters.

The reference keeps one ``nn.Parameter`` per small matrix (3 501 tensors for the 500-stage SSS layer,
``layers/sss_layer.py:83-89``; two per leaf for the H-matrix layer, ``hmatrix/hmatrix_component.py:13-18``).
The kernels, the optimizer and the data-parallel all-reduce all want ONE contiguous buffer, so the
modules here keep every float32 parameter as a *view* into a single flat tensor (same names, shapes
and ``state_dict`` keys as the reference) and every ``.grad`` as a view into a single flat gradient
tensor that the backward kernels accumulate into directly (SURVEY.md section 7 "3 501 parameter tensors").
"""
import copy
from typing import List

import torch
import torch.nn as nn


def _strip_params(params, flat_ids):
    return type(params)((k, (None if (p is not None and id(p) in flat_ids) else p)) for k, p in params.items())


def _strip_module(module, flat_ids):
    if module is None:
        return None
    clone = copy.copy(module)
    clone._parameters = _strip_params(module._parameters, flat_ids)
    clone._modules = type(module._modules)((k, _strip_module(m, flat_ids)) for k, m in module._modules.items())
    return clone


class FlatParamsMixin:
    """Mixin for nn.Modules.  Subclasses call ``_flatten_parameters()`` at the end of ``__init__``."""

    # ---- construction ------------------------------------------------------------------------
    def _flat_param_list(self) -> List[nn.Parameter]:
        """Parameters that live in the flat buffer, in ``named_parameters()`` order (cached: walking
        thousands of ParameterList entries costs more than a kernel launch)."""
        cached = self.__dict__.get("_flat_params_cache")
        if cached is None:
            cached = [p for p in self.parameters() if p.dtype == torch.float32 and not p.is_sparse]
            self.__dict__["_flat_params_cache"] = cached
        return cached

    def _flatten_parameters(self):
        self.__dict__["_flat_params_cache"] = None
        params = self._flat_param_list()
        total = sum(p.numel() for p in params)
        device = params[0].device if params else torch.device("cpu")
        flat = torch.empty(max(total, 1), dtype=torch.float32, device=device)
        offsets = []
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                view = flat[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                offsets.append(off)
                off += n
        self.__dict__["_flat"] = flat
        self.__dict__["_flat_offsets"] = offsets
        self.__dict__["_flat_total"] = total
        self.__dict__["_flat_grad"] = None
        self.__dict__["_flat_grad_views"] = None
        self.__dict__["_flat_sentinels"] = self._make_sentinels(params)
        self._on_reflatten()

    @staticmethod
    def _make_sentinels(params):
        if not params:
            return ()
        return (params[0].data_ptr(), params[-1].data_ptr())

    def _on_reflatten(self):
        """Hook: device-side plans cached by subclasses must be dropped here."""

    def _flat_is_valid(self) -> bool:
        flat = self.__dict__.get("_flat")
        if flat is None:
            return False
        params = self._flat_param_list()
        return self._make_sentinels(params) == self.__dict__.get("_flat_sentinels") and \
            (not params or params[0].device == flat.device)

    def _ensure_flat(self):
        if not self._flat_is_valid():
            self._flatten_parameters()

    # ---- nn.Module plumbing ------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.float() replace every parameter's storage separately; re-pack afterwards
        res = super()._apply(fn, *args, **kwargs)
        if "_flat" in self.__dict__:
            self._flatten_parameters()
        return res

    def __getstate__(self):
        # A pickled view-parameter would carry the whole flat storage (torch pickles storages per
        # tensor), i.e. thousands of copies.  Pickle the flat buffer once plus a manifest, and strip
        # the view parameters out of the module tree; __setstate__ re-creates them as views.
        self._ensure_flat()
        state = self.__dict__.copy()
        for k in ("_flat_grad", "_flat_grad_views", "_flat_sentinels", "_flat_params_cache"):
            state.pop(k, None)
        for k in list(state.keys()):
            if k.startswith("_dev_"):
                state.pop(k)
        flat_ids = {id(p): i for i, p in enumerate(self._flat_param_list())}
        manifest = []
        for name, p in self.named_parameters():
            if id(p) in flat_ids:
                manifest.append((name, tuple(p.shape), self.__dict__["_flat_offsets"][flat_ids[id(p)]], p.requires_grad))
        state["_flat_manifest"] = manifest
        state["_parameters"] = _strip_params(self._parameters, flat_ids)
        state["_modules"] = type(self._modules)((k, _strip_module(m, flat_ids)) for k, m in self._modules.items())
        return state

    def __setstate__(self, state):
        state = dict(state)
        manifest = state.pop("_flat_manifest", [])
        super().__setstate__(state)
        flat = self.__dict__["_flat"]
        for name, shape, off, rg in manifest:
            owner = self
            parts = name.split(".")
            for part in parts[:-1]:
                owner = owner._modules[part]
            n = 1
            for d in shape:
                n *= d
            owner._parameters[parts[-1]] = nn.Parameter(flat[off:off + n].view(shape), requires_grad=rg)
        self.__dict__["_flat_grad"] = None
        self.__dict__["_flat_grad_views"] = None
        self.__dict__["_flat_params_cache"] = None
        self.__dict__["_flat_sentinels"] = self._make_sentinels(self._flat_param_list())
        self._on_reflatten()

    # ---- gradients ---------------------------------------------------------------------------
    def flat_parameters(self) -> torch.Tensor:
        """The contiguous float32 buffer all dense parameters are views of."""
        self._ensure_flat()
        return self.__dict__["_flat"]

    def flat_grad(self) -> torch.Tensor:
        """The contiguous float32 gradient buffer (allocated on first use, zero-filled)."""
        self._ensure_flat()
        g = self.__dict__.get("_flat_grad")
        flat = self.__dict__["_flat"]
        if g is None or g.device != flat.device or g.numel() != flat.numel():
            g = torch.zeros_like(flat)
            self.__dict__["_flat_grad"] = g
            views = []
            for p, off in zip(self._flat_param_list(), self.__dict__["_flat_offsets"]):
                views.append(g[off:off + p.numel()].view(p.shape))
            self.__dict__["_flat_grad_views"] = views
        return g

    def zero_flat_grad(self):
        """One memset instead of one per parameter tensor; keeps every ``p.grad`` bound."""
        g = self.__dict__.get("_flat_grad")
        if g is not None:
            g.zero_()

    def _prepare_grad_accumulation(self) -> torch.Tensor:
        """Called by the backward of the layer's autograd Function, before the kernels accumulate.

        If the parameters' ``.grad`` are unset (fresh module or ``optimizer.zero_grad()`` with
        ``set_to_none=True``) the flat buffer is cleared and every ``.grad`` is (re)bound to its view;
        otherwise the kernels accumulate on top of what is there, like autograd does.
        """
        g = self.flat_grad()
        params = self._flat_param_list()
        views = self.__dict__["_flat_grad_views"]
        # sentinels: the first and the last parameter that TAKE a gradient (a frozen parameter's .grad stays None for ever; using it
        # as a sentinel would clear the buffer in every backward and lose gradient accumulation)
        # (looked up once: walking the 3 501 parameters of an SSS layer in every backward cost 150 us of host time per step; the cached
        # pair stays valid as long as both still take gradients)
        live = self.__dict__.get("_flat_live_sentinels")
        if live is None or len(live) != 2 or live[1] >= len(params) or not (params[live[0]].requires_grad and params[live[1]].requires_grad):
            idx = [i for i, p in enumerate(params) if p.requires_grad]
            live = (idx[0], idx[-1]) if idx else ()
            self.__dict__["_flat_live_sentinels"] = live
        if live and (params[live[0]].grad is None or params[live[-1]].grad is None
                     or params[live[0]].grad.data_ptr() != views[live[0]].data_ptr()):
            g.zero_()
            for p, v in zip(params, views):
                if p.requires_grad:
                    p.grad = v
        return g
