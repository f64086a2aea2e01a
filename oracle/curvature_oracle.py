"""CPU oracle for the curvature-matvec hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``curvlinops_b200``) never does, and fails loudly when its CUDA library is missing.

What it restates (all ``file:line`` relative to ``/root/reference``):

* the data loop / normalisation of ``CurvatureLinearOperator._matmat``
  (``curvlinops/_torch_base.py:923-944``, ``curvlinops/_empirical_risk.py:340-352``),
* the three-step GGN product ``J^T (H_loss (J v))`` of ``make_ggn_vector_product``
  (``curvlinops/ggn.py:42-72``) with *explicit* loss-Hessian formulas taken from the
  documented square roots in ``curvlinops/ggn_utils.py:29-171``,
* the Monte-Carlo pseudo-loss of ``make_batch_ggn_mc_vector_product``
  (``curvlinops/ggn.py:100-168``, sampler ``curvlinops/ggn_utils.py:174-271``),
* the Hessian product of ``make_batch_hessian_vector_product`` (``curvlinops/hessian.py:13-69``),
* the KFAC factor conventions of the hooks backend
  (``curvlinops/computers/kfac_hooks.py:176-393``, ``curvlinops/computers/kfac_math.py:47-203``)
  and the Kronecker damped-inverse apply (``curvlinops/kronecker.py:141-171,250-373``).

How it differs from the reference (so that it is an independent restatement, not a copy):
the model Jacobian is applied with plain reverse-mode autograd only -- ``J^T w`` is
``autograd.grad`` and ``J v`` is the derivative of ``<J^T w, v>`` w.r.t. the dummy ``w``
(double-vjp trick) -- instead of ``torch.func.jvp/vjp/jacrev/vmap``; the loss Hessian is
applied in closed form; the columns of ``V`` are looped.

Parity pin: ``tests/golden/*.npz`` were produced by the *reference itself* in the build
container (``oracle/make_golden.py``); ``tests/test_oracle_golden.py`` checks this module
against them, so the oracle is pinned.
"""

from __future__ import annotations

import math
from typing import Callable, Iterable

import torch
from torch import Tensor
from torch.nn import BCEWithLogitsLoss, CrossEntropyLoss, MSELoss
from torch.func import functional_call


def _as_callable(model_func) -> Callable:
    """``(params, X) -> prediction``; nn.Modules are called with overridden parameters
    (reference ``curvlinops/utils.py:267-297``)."""
    if isinstance(model_func, torch.nn.Module):
        return lambda p, X: functional_call(model_func, p, (X,))
    return model_func


def _normalization(loss_func, batch_size: int, n_data: int) -> float:
    """Per-batch weight (reference ``curvlinops/_empirical_risk.py:340-352``)."""
    return {"sum": 1.0, "mean": batch_size / n_data}[loss_func.reduction]


def loss_hessian_apply(loss_func, f: Tensor, y: Tensor, u: Tensor) -> Tensor:
    """``(nabla_f^2 loss)(f, y) @ u`` for a 2-d prediction ``f [B, C]``, reduction included.

    CE: ``c (diag(p) - p p^T)`` per datum with ``c = 1/B`` (mean) or 1 (sum);
    MSE: ``2 c I`` with ``c = 1/(B C)`` or 1; BCE: ``c diag(s (1-s))``, same ``c`` as MSE
    (``curvlinops/ggn_utils.py:57-108`` for the per-datum factors; the extra ``1/B`` is the
    batch mean of ``torch.nn`` losses).
    """
    B = f.shape[0]
    if isinstance(loss_func, CrossEntropyLoss):
        c = 1.0 / B if loss_func.reduction == "mean" else 1.0
        p = torch.softmax(f, dim=1)
        return c * (p * u - p * (p * u).sum(dim=1, keepdim=True))
    if isinstance(loss_func, MSELoss):
        c = 1.0 / f.numel() if loss_func.reduction == "mean" else 1.0
        return 2.0 * c * u
    if isinstance(loss_func, BCEWithLogitsLoss):
        c = 1.0 / f.numel() if loss_func.reduction == "mean" else 1.0
        s = torch.sigmoid(f)
        return c * s * (1.0 - s) * u
    raise NotImplementedError(f"Unsupported loss {loss_func}.")


def mc_grad_outputs(loss_func, f: Tensor, mc_samples: int) -> Tensor:
    """Sampled would-be gradients ``[B, M, C]`` scaled by ``1/sqrt(M)``, drawn from the
    *global* torch RNG exactly like the reference does inside ``fork_rng`` after
    ``manual_seed(seed)`` (``curvlinops/ggn.py:337-341``, ``curvlinops/ggn_utils.py:220-262,369-372``).
    Per-datum factor ``sqrt(c)`` with ``c = 1`` (CE, sum) / ``1/C`` (MSE/BCE mean)."""
    B, C = f.shape
    M = mc_samples
    if isinstance(loss_func, CrossEntropyLoss):
        p = torch.softmax(f, dim=1)
        # vmap(randomness="different") over the batch == one batched multinomial call
        yhat = p.multinomial(M, replacement=True)  # [B, M]
        onehot = torch.nn.functional.one_hot(yhat, num_classes=C).to(f.dtype)
        g = p.unsqueeze(1) - onehot
    elif isinstance(loss_func, MSELoss):
        c = 1.0 / C if loss_func.reduction == "mean" else 1.0
        g = torch.normal(torch.zeros(B, M, C, dtype=f.dtype), math.sqrt(2 * c))
    elif isinstance(loss_func, BCEWithLogitsLoss):
        c = 1.0 / C if loss_func.reduction == "mean" else 1.0
        s = torch.sigmoid(f).unsqueeze(1).expand(B, M, C)
        g = math.sqrt(c) * (s - s.bernoulli())
    else:
        raise NotImplementedError(f"Unsupported loss {loss_func}.")
    return g / math.sqrt(M)


def _jt_w(f_fn: Callable, params: dict, X, w: Tensor, create_graph=False):
    pred = f_fn(params, X)
    return pred, torch.autograd.grad(pred, list(params.values()), w, create_graph=create_graph,
                                     allow_unused=True)


def jacobian_vector_product(f_fn, params: dict, X, v: list[Tensor]) -> tuple[Tensor, Tensor]:
    """``(f, J v)`` through the double-vjp trick (reference uses ``torch.func.jvp``,
    ``curvlinops/ggn.py:61``, ``curvlinops/jacobian.py:47``)."""
    p = {k: t.detach().requires_grad_(True) for k, t in params.items()}
    pred = f_fn(p, X)
    w = torch.zeros_like(pred, requires_grad=True)
    g = torch.autograd.grad(pred, list(p.values()), w, create_graph=True, allow_unused=True)
    s = sum((gi * vi).sum() for gi, vi in zip(g, v) if gi is not None)
    (jv,) = torch.autograd.grad(s, w)
    return pred.detach(), jv.detach()


def transposed_jacobian_vector_product(f_fn, params: dict, X, w: Tensor) -> list[Tensor]:
    """``J^T w`` (reference ``curvlinops/ggn.py:70-71``, ``curvlinops/jacobian.py:92``)."""
    p = {k: t.detach().requires_grad_(True) for k, t in params.items()}
    pred = f_fn(p, X)
    g = torch.autograd.grad(pred, list(p.values()), w, allow_unused=True)
    return [torch.zeros_like(t) if gi is None else gi.detach() for gi, t in zip(g, p.values())]


def ggn_matmat(model_func, loss_func, params: dict[str, Tensor],
               data: Iterable[tuple[Tensor, Tensor]], V: list[Tensor], n_data: int | None = None,
               mc_samples: int = 0, seed: int = 2147483647) -> list[Tensor]:
    """``GGN @ V`` with ``V`` in tensor-list format ``[*shape_i, K]``.

    Loop over mini-batches with weight ``B_b/N`` or 1 (``_torch_base.py:937-944``); per
    batch and column: ``Jv`` -> loss Hessian (exact, or rank-M MC estimate) -> ``J^T``.
    """
    f_fn = _as_callable(model_func)
    data = list(data)
    if n_data is None:
        n_data = sum(X.shape[0] for X, _ in data)
    K = V[0].shape[-1]
    out = [torch.zeros_like(v) for v in V]

    def run():
        for X, y in data:
            alpha = _normalization(loss_func, X.shape[0], n_data)
            g_mc = None
            for k in range(K):
                vk = [v[..., k] for v in V]
                f, jv = jacobian_vector_product(f_fn, params, X, vk)
                if mc_samples > 0:
                    if g_mc is None:  # same samples for all columns (vmap randomness="same")
                        g_mc = mc_grad_outputs(loss_func, f, mc_samples)
                    red = f.shape[0] if loss_func.reduction == "mean" else 1.0
                    ip = torch.einsum("nmc,nc->nm", g_mc, jv)
                    hjv = torch.einsum("nmc,nm->nc", g_mc, ip) / red
                else:
                    hjv = loss_hessian_apply(loss_func, f, y, jv)
                for o, g in zip(out, transposed_jacobian_vector_product(f_fn, params, X, hjv)):
                    o[..., k].add_(g, alpha=alpha)

    if mc_samples > 0:
        with torch.random.fork_rng():
            torch.manual_seed(seed)
            run()
    else:
        run()
    return out



def loss_grad_outputs(loss_func, f: Tensor, y: Tensor) -> Tensor:
    """Per-sample gradient of the UNREDUCED loss w.r.t. the prediction, ``g_n = d l(f_n, y_n) / d f_n`` -- what
    ``grad(c_flat)(output) * reduction_factor`` evaluates to in the reference's empirical-Fisher pseudo-loss
    (``curvlinops/gradient_moments.py:65-80``)."""
    if isinstance(loss_func, torch.nn.CrossEntropyLoss):
        return torch.softmax(f, dim=1) - torch.nn.functional.one_hot(y, f.shape[1]).to(f.dtype)
    if isinstance(loss_func, torch.nn.MSELoss):
        return 2.0 * (f - y)
    if isinstance(loss_func, torch.nn.BCEWithLogitsLoss):
        return torch.sigmoid(f) - y
    raise NotImplementedError(f"loss {loss_func}")


def ef_matmat(model_func, loss_func, params: dict[str, Tensor], data: Iterable[tuple[Tensor, Tensor]],
              V: list[Tensor], n_data: int | None = None) -> list[Tensor]:
    """``EF @ V`` (uncentered gradient covariance, ``curvlinops/gradient_moments.py:15-151``): the GGN of the
    pseudo-loss ``1/(2c) sum_n <f_n, g_n>^2``, i.e. the two sweeps with the rank-one per-sample "loss Hessian"
    ``g_n g_n^T / c`` (c = number of loss terms of a mean reduction, 1 for sum)."""
    f_fn = _as_callable(model_func)
    data = list(data)
    if n_data is None:
        n_data = sum(X.shape[0] for X, _ in data)
    K = V[0].shape[-1]
    out = [torch.zeros_like(v) for v in V]
    for X, y in data:
        alpha = _normalization(loss_func, X.shape[0], n_data)
        for k in range(K):
            vk = [v[..., k] for v in V]
            f, jv = jacobian_vector_product(f_fn, params, X, vk)
            g = loss_grad_outputs(loss_func, f.detach(), y)
            if loss_func.reduction == "mean":
                red = f.shape[0] if isinstance(loss_func, torch.nn.CrossEntropyLoss) else f.numel()
            else:
                red = 1.0
            hjv = g * (g * jv).sum(dim=1, keepdim=True) / red
            for o, gr in zip(out, transposed_jacobian_vector_product(f_fn, params, X, hjv)):
                o[..., k].add_(gr, alpha=alpha)
    return out

def hessian_matmat(model_func, loss_func, params: dict[str, Tensor],
                   data: Iterable[tuple[Tensor, Tensor]], V: list[Tensor],
                   n_data: int | None = None) -> list[Tensor]:
    """``Hessian @ V`` by reverse-over-reverse (the reference does forward-over-reverse,
    ``curvlinops/hessian.py:66``)."""
    f_fn = _as_callable(model_func)
    data = list(data)
    if n_data is None:
        n_data = sum(X.shape[0] for X, _ in data)
    K = V[0].shape[-1]
    out = [torch.zeros_like(v) for v in V]
    for X, y in data:
        alpha = _normalization(loss_func, X.shape[0], n_data)
        p = {k: t.detach().requires_grad_(True) for k, t in params.items()}
        loss = loss_func(f_fn(p, X), y)
        g = torch.autograd.grad(loss, list(p.values()), create_graph=True, allow_unused=True)
        for k in range(K):
            s = sum((gi * v[..., k]).sum() for gi, v in zip(g, V) if gi is not None)
            hv = torch.autograd.grad(s, list(p.values()), retain_graph=True, allow_unused=True)
            for o, h in zip(out, hv):
                if h is not None:
                    o[..., k].add_(h.detach(), alpha=alpha)
    return out


def gradient_and_loss(model_func, loss_func, params, data, n_data=None):
    """Total gradient and loss (reference ``curvlinops/_empirical_risk.py:409-439``)."""
    f_fn = _as_callable(model_func)
    data = list(data)
    if n_data is None:
        n_data = sum(X.shape[0] for X, _ in data)
    tot = [torch.zeros_like(t) for t in params.values()]
    tot_loss = 0.0
    for X, y in data:
        alpha = _normalization(loss_func, X.shape[0], n_data)
        p = {k: t.detach().requires_grad_(True) for k, t in params.items()}
        loss = loss_func(f_fn(p, X), y) * alpha
        g = torch.autograd.grad(loss, list(p.values()), allow_unused=True)
        tot_loss += float(loss)
        for t, gi in zip(tot, g):
            if gi is not None:
                t.add_(gi)
    return tot, tot_loss


# ----------------------------------------------------------------------------------------
# KFAC (hooks-backend conventions)
# ----------------------------------------------------------------------------------------
def _patches(x: Tensor, mod: torch.nn.Conv2d) -> Tensor:
    """``[B, S, C_in*k*k]`` im2col patches, channel-major like ``F.unfold``
    (reference ``curvlinops/kfac_utils.py:78-121``)."""
    u = torch.nn.functional.unfold(x, mod.kernel_size, dilation=mod.dilation,
                                   padding=mod.padding, stride=mod.stride)
    return u.transpose(1, 2)


def kfac_factors(model: torch.nn.Module, loss_func, layer_names: list[str],
                 data, n_data: int | None = None, fisher_type: str = "mc", mc_samples: int = 1,
                 seed: int = 2147483647, joint_bias: bool = True,
                 grad_outputs: list[Tensor] | None = None):
    """KFAC-expand Kronecker factors ``(A_l, G_l)`` for Linear/Conv2d layers.

    ``A = sum a~ a~^T / (N S)`` with a ones column for a joint bias
    (``kfac_hooks.py:355-393``, ``kfac_math.py:47-118``); ``G = sum g g^T / (T N)`` for
    mean reduction, where g are the backpropagated would-be gradients
    (``kfac_hooks.py:236-353``, ``kfac_math.py:172-203``); T = per-example loss terms = 1 here.

    ``fisher_type``: ``"type2"`` (columns of the loss-Hessian square root),
    ``"mc"`` (draws with a dedicated ``torch.Generator(seed)``, ``kfac_hooks.py:219``),
    ``"empirical"``.  ``grad_outputs`` (one ``[V, B, C]`` tensor per batch, *unscaled*
    per-datum vectors) overrides the draw -- that is how the engine is handed identical
    samples in the parity tests.
    """
    data = list(data)
    if n_data is None:
        n_data = sum(X.shape[0] for X, _ in data)
    mods = dict(model.named_modules())
    A = {n: None for n in layer_names}
    G = {n: None for n in layer_names}
    gen = torch.Generator()
    gen.manual_seed(seed)
    for bi, (X, y) in enumerate(data):
        store_in, store_out, hooks = {}, {}, []
        def make_hook(n):
            def hook(m, i, o):
                store_in[n] = i[0].detach()
                store_out[n] = o
            return hook

        for n in layer_names:
            hooks.append(mods[n].register_forward_hook(make_hook(n)))
        out = model(X)
        for h in hooks:
            h.remove()
        B, C = out.shape
        f = out.detach()
        if grad_outputs is not None:
            gos = grad_outputs[bi]
        elif fisher_type == "type2":
            if isinstance(loss_func, CrossEntropyLoss):
                p = torch.softmax(f, 1)
                S = torch.diag_embed(p.sqrt()) - p.unsqueeze(2) * p.sqrt().unsqueeze(1)  # [B,C,C]
                gos = S.permute(2, 0, 1)  # [V=C, B, C]
            elif isinstance(loss_func, MSELoss):
                c = 1.0 / C if loss_func.reduction == "mean" else 1.0
                gos = math.sqrt(2 * c) * torch.eye(C, dtype=f.dtype).unsqueeze(1).expand(C, B, C)
            else:
                c = 1.0 / C if loss_func.reduction == "mean" else 1.0
                s = torch.sigmoid(f)
                gos = math.sqrt(c) * torch.diag_embed((s * (1 - s)).sqrt()).permute(2, 0, 1)
        elif fisher_type == "mc":
            if isinstance(loss_func, CrossEntropyLoss):
                p = torch.softmax(f, 1)
                # the reference vmaps a per-datum sampler over the batch with a generator
                yhat = torch.stack([p[b:b + 1].multinomial(mc_samples, replacement=True,
                                                           generator=gen)[0] for b in range(B)])
                gos = (p.unsqueeze(1) - torch.nn.functional.one_hot(yhat, C).to(f.dtype))
                gos = gos.permute(1, 0, 2) / math.sqrt(mc_samples)
            else:
                raise NotImplementedError
        else:
            raise NotImplementedError(fisher_type)
        V = gos.shape[0]
        scale = 1.0 / B if loss_func.reduction == "mean" else 1.0  # seeds scaled 1/(B T)
        corr = (B * B / n_data) if loss_func.reduction == "mean" else 1.0  # (B T)^2/(T N)
        for n in layer_names:
            m, x = mods[n], store_in[n]
            if isinstance(m, torch.nn.Conv2d):
                a = _patches(x, m)
            else:
                a = x.reshape(B, -1, x.shape[-1])
            S_pos = a.shape[1]
            if joint_bias and m.bias is not None:
                a = torch.cat([a, a.new_ones(*a.shape[:-1], 1)], dim=-1)
            cur = torch.einsum("bsi,bsj->ij", a, a) / (n_data * S_pos)
            A[n] = cur if A[n] is None else A[n] + cur
        for v in range(V):
            gs = torch.autograd.grad(out, [store_out[n] for n in layer_names],
                                     gos[v] * scale, retain_graph=True)
            for n, g in zip(layer_names, gs):
                m = mods[n]
                g2 = g.flatten(2).transpose(1, 2) if isinstance(m, torch.nn.Conv2d) \
                    else g.reshape(B, -1, g.shape[-1])
                cur = torch.einsum("bsi,bsj->ij", g2, g2) * corr
                G[n] = cur if G[n] is None else G[n] + cur
    return A, G


def kfac_grad_outputs(model, loss_func, data, fisher_type="mc", mc_samples=1, seed=2147483647):
    """The unscaled per-datum seed vectors ``[V, B, C]`` per mini-batch that :func:`kfac_factors` back-propagates
    (same draws: dedicated generator, per-datum multinomial).  Used to hand the engine identical samples."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    res = []
    for X, y in data:
        f = model(X).detach()
        B, C = f.shape
        if fisher_type == "mc":
            p = torch.softmax(f, 1)
            yhat = torch.stack([p[b:b + 1].multinomial(mc_samples, replacement=True, generator=gen)[0]
                                for b in range(B)])
            g = p.unsqueeze(1) - torch.nn.functional.one_hot(yhat, C).to(f.dtype)
            res.append(g.permute(1, 0, 2) / math.sqrt(mc_samples))
        elif fisher_type == "empirical":
            ff = f.clone().requires_grad_(True)
            (g,) = torch.autograd.grad(type(loss_func)(reduction="sum")(ff, y), ff)
            res.append(g.unsqueeze(0))
        else:
            raise NotImplementedError(fisher_type)
    return res


def loss_hessian_sqrt_columns(loss_func, f: Tensor) -> Tensor:
    """``[C', B, C]``: per datum the columns of ``S`` with ``S S^T`` = Hessian of the loss w.r.t. the prediction, without
    the 1/B of a mean reduction (reference ``curvlinops/ggn_utils.py:29-171``)."""
    B, C = f.shape
    c = 1.0 / C if (loss_func.reduction == "mean" and not isinstance(loss_func, torch.nn.CrossEntropyLoss)) else 1.0
    if isinstance(loss_func, torch.nn.CrossEntropyLoss):
        p = torch.softmax(f, 1)
        S = torch.diag_embed(p.sqrt()) - p.unsqueeze(2) * p.sqrt().unsqueeze(1)  # [B, C, C']
    elif isinstance(loss_func, torch.nn.MSELoss):
        S = (math.sqrt(2 * c) * torch.eye(C, dtype=f.dtype, device=f.device)).expand(B, C, C)
    elif isinstance(loss_func, torch.nn.BCEWithLogitsLoss):
        s = torch.sigmoid(f)
        S = math.sqrt(c) * torch.diag_embed((s * (1 - s)).sqrt())
    else:
        raise NotImplementedError(type(loss_func))
    return S.permute(2, 0, 1).contiguous()


def ggn_diagonal(model_func, loss_func, params: dict[str, Tensor], data: Iterable[tuple[Tensor, Tensor]],
                 grad_outputs: list[Tensor] | None = None, n_data: int | None = None) -> list[Tensor]:
    """Diagonal of the GGN, one tensor per parameter (reference ``curvlinops/computers/ggn_diagonal.py:49-109,207-232``):
    ``w * sum_{n, v} (J_n^T g'_{n,v})^2`` with ``g'`` the loss-Hessian square-root columns (exact) or the given
    ``grad_outputs`` (one ``[V, B, C]`` tensor per mini-batch, e.g. MC samples), ``w = 1/N`` (mean) or 1 (sum).
    Plain loops over data points and columns: small cases only."""
    f_fn = _as_callable(model_func)
    data = list(data)
    N = n_data if n_data is not None else sum(X.shape[0] for X, _ in data)
    ps = list(params.values())
    out = [torch.zeros_like(p) for p in ps]
    w = 1.0 / N if loss_func.reduction == "mean" else 1.0
    for bi, (X, y) in enumerate(data):
        with torch.enable_grad():
            f = f_fn(params, X)
            seeds = loss_hessian_sqrt_columns(loss_func, f.detach()) if grad_outputs is None else grad_outputs[bi]
            for n in range(X.shape[0]):
                for v in range(seeds.shape[0]):
                    if not bool(seeds[v, n].any()):
                        continue
                    gs = torch.autograd.grad(f[n], ps, grad_outputs=seeds[v, n].to(f.dtype), retain_graph=True,
                                             allow_unused=True)
                    for o, g in zip(out, gs):
                        if g is not None:
                            o.add_(g.detach() ** 2, alpha=w)
    return out


def damped_inverse(S: Tensor, damping: float) -> Tensor:
    """``(S + damping I)^-1`` via Cholesky (reference ``curvlinops/kronecker.py:328-373``)."""
    L = torch.linalg.cholesky(S + damping * torch.eye(S.shape[0], dtype=S.dtype))
    return torch.cholesky_inverse(L)


def kron_apply(Gm: Tensor, Am: Tensor, W: Tensor) -> Tensor:
    """``(G (x) A) vec(W)`` for ``W [d_out, d_in, K]`` = ``G W A^T`` per column
    (reference ``curvlinops/kronecker.py:141-153``)."""
    return torch.einsum("abz,Aa,Bb->ABz", W, Gm, Am)
