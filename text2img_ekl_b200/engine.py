"""The G+D training step (hot loop of cub_trainer_splitz_cap_ca.py:547-608 / trainer.py:509-545), B200-native.

What changes against the reference, with semantics preserved:
  * the three discriminator forwards of a D update (real / wrong / fake, cub:418-420) run as ONE pass over the
    stacked batch with per-group BatchNorm statistics (weights are read once instead of three times);
  * gradients accumulate in place into one flat fp32 buffer per network (zeroed with a single memset; this is
    also the NCCL all-reduce bucket when WORLD_SIZE > 1, replacing nn.DataParallel's reduce-to-GPU0);
  * the discriminator weight gradients that `errGs.backward()` produces in the reference and the next
    `netD.zero_grad()` throws away (cub:414,607) are not computed;
  * label tensors (one-hot / multi-hot normalisation) are built on the device.
RNG draws (noise, CA eps, VC seed) can be injected for parity tests; otherwise they are drawn on the device.
"""
import os

import torch
import torch.nn.functional as F

from . import ops
from .miscc.config import cfg


def KL_loss(mu, logvar):
    """cub:54-58: -0.5 * mean(1 + logvar - mu^2 - exp(logvar))."""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def ce_loss(logq, p, average=True):
    """cub:60-65: -sum(p * logq) / B."""
    n = p.shape[0] if average else 1
    return -torch.sum(p * logq) / n


def compute_mean_covariance(img):
    """cub:33-52: per-image channel mean [B,C,1,1] and channel covariance [B,C,C] (colour-consistency statistics)."""
    if img.is_cuda:
        x = img.float().contiguous()
        if not ops.color_stats_supported(x):
            raise ops.L.EklError("compute_mean_covariance on the GPU takes [B,3,H,W] images with H*W %% 4 == 0, got %s"
                                 % (tuple(img.shape),))
        return ops.color_stats(x)             # one pass over the images (include/ekl_b200.h: ekl_color_stats_fwd)
    b, c, h, w = img.shape                    # host tensors (tests, tooling): the reference formula
    mu = img.mean(2, keepdim=True).mean(3, keepdim=True)
    d = (img - mu).reshape(b, c, h * w)
    return mu, torch.bmm(d, d.transpose(1, 2)) / (h * w)


def onehot(cls_vec, n):
    """cub:322-331 on the device: [B] int64 class index -> [B,n] float one-hot."""
    out = torch.zeros(cls_vec.shape[0], n, device=cls_vec.device)
    out.scatter_(1, cls_vec.view(-1, 1).long(), 1.0)
    return out


def _bce_const(p, target):
    """nn.BCELoss() against a constant 0/1 label vector (cub:423-431); log clamped at -100 like torch."""
    if target:
        return -torch.clamp(torch.log(p), min=-100.0).mean()
    return -torch.clamp(torch.log1p(-p), min=-100.0).mean()


class tensor_core_matmul:
    """The path's few dense Linear layers (CA_NET / VC_NET, the generator stem, the class head fc_ac) are plain library
    GEMMs on fp32 parameters.  With TF32 off cuBLAS runs them as SIMT sgemm (0.8 ms of an 17 ms step on B200);
    inside the step they run as TF32 tensor-core GEMMs (fp32 accumulate), which is within the north star's stated
    BF16/TF32 tolerance.  The caller's global flag is restored on exit."""

    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


class FlatGrads:
    """All gradients of a network as views into one flat fp32 buffer (same memory layout as each parameter)."""

    def __init__(self, params, optimizer=None):
        self.params = [p for p in params if p.requires_grad]
        if hasattr(optimizer, "make_flat_grads"):        # optim.FlatAdam: gradients share the parameters' flat offsets
            self.flat = optimizer.make_flat_grads()
            self.offsets = list(optimizer.offsets)
            return
        pad4 = lambda n: (n + 3) // 4 * 4                 # 16-byte aligned slices (the gradient reducer's granularity)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += pad4(p.numel())
        dev = self.params[0].device
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        for p, off in zip(self.params, self.offsets):
            g = self.flat[off:off + p.numel()]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                g = g.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
            else:
                g = g.view(p.shape)
            p.grad = g

    def zero(self):
        self.flat.zero_()


class TailUpdate:
    """Optimiser step of ONE discriminator overlapped with its own backward pass.

    The parameters sit in the flat buffers in registration = forward order, so the deepest layers -- whose gradients
    backward produces first and which hold most of the weights (JOINT_D_NET256: the 1024->2048 4x4 and 2048->1024 3x3
    filters are 52 M of 73 M parameters) -- form the tail.  The network's forward marks block boundaries
    (ops.grad_mark); when the gradient of a marked activation arrives every parameter registered after that block is
    final (and the backward no longer reads those filters), so the Adam update of that slice runs on a side stream while
    the main stream continues the backward through the high-resolution layers.  finish() updates the remaining head slice
    and joins.  Adam is element-wise: the pieces give bit-identical results to one full step.

    With several ranks the slices are the gradient reducer's (parallel.GradReducer): each slice is updated on the
    reducer's side stream right behind its all-reduce."""

    MIN_ELEMS = 1 << 20

    def __init__(self, opt, net, grads, reducer=None):
        self.opt, self.red, self.total = opt, reducer, opt.n
        self.side = torch.cuda.Stream() if reducer is None else None
        self.end_of = {}
        end_of_param = {id(p): o + (p.numel() + 7) // 8 * 8 for p, o in zip(grads.params, grads.offsets)}
        for m in net.modules():
            ends = [end_of_param[id(p)] for p in m.parameters() if id(p) in end_of_param]
            if ends:
                self.end_of[id(m)] = min(max(ends), self.total)
        self.active, self.lo = False, self.total
        if reducer is not None:
            reducer.after_slice = self._after_reduced

    def begin(self):
        """Right before backward of the discriminator's own update (marks fired at any other time are ignored)."""
        self.opt.tick()
        self.active, self.lo, self.forked = True, self.total, False
        if self.red is not None:
            self.red.begin()

    def _after_reduced(self, lo, hi):
        # on the reducer's side stream, behind the all-reduce of [lo, hi)
        self.opt.apply_slice(lo, hi, self.red.stage)

    def on_mark(self, after_module):
        if not self.active:
            return
        if self.red is not None:
            return self.red.on_mark(after_module)
        lo = self.end_of.get(id(after_module))
        if lo is None or self.lo - lo < self.MIN_ELEMS:
            return
        self.side.wait_stream(torch.cuda.current_stream())      # the slice's gradients are complete on the issuing stream
        with torch.cuda.stream(self.side):
            self.opt.apply_slice(lo, self.lo)
        self.lo, self.forked = lo, True

    def finish(self):
        self.active = False
        if self.red is not None:
            self.red.finish()               # sends the head slice (updated behind its all-reduce) and joins the side stream
        else:
            self.opt.apply_slice(0, self.lo)
            if self.forked:
                torch.cuda.current_stream().wait_stream(self.side)
            self.lo = 0
        self.opt.finish_step()


class BnCounters:
    """nn.BatchNorm's num_batches_tracked bookkeeping (one += per forward call in the reference) for a whole network
    as ONE kernel per step: every counter becomes a 0-dim view into one int64 tensor (state_dict keys / values are
    unchanged), the per-call increments are tallied on the host and added in a single vector add."""

    def __init__(self, nets):
        self.mods = [m for net in nets for m in net.modules()
                     if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None]
        if not self.mods:
            self.flat = None
            return
        dev = self.mods[0].num_batches_tracked.device
        self.flat = torch.stack([m.num_batches_tracked.detach().to(dev).long().reshape(()) for m in self.mods])
        for i, m in enumerate(self.mods):
            m._buffers["num_batches_tracked"] = self.flat[i]
            m._ekl_counted, m._ekl_calls = True, 0
        self._inc_key, self._inc = None, None

    def flush(self):
        if self.flat is None:
            return
        key = tuple(m._ekl_calls for m in self.mods)
        if not any(key):
            return
        if key != self._inc_key:      # host->device copy: happens on the first step(s), before any graph capture
            self._inc_key, self._inc = key, torch.tensor(key, dtype=torch.long, device=self.flat.device)
        self.flat.add_(self._inc)
        for m in self.mods:
            m._ekl_calls = 0


class StepEngine:
    """kind 'catz_ca': COND_G_NET_CATZ_CA flavour (cub trainer); 'cond': COND_G_NET flavour (trainer.py)."""

    def __init__(self, netG, netsD, optimizerG, optimizersD, kind, cond="txt+cls", data_parallel=None):
        self.netG, self.netsD = netG, netsD
        self.optG, self.optsD = optimizerG, optimizersD
        self.kind, self.cond = kind, cond
        self.gradsG = FlatGrads(netG.parameters(), optimizerG)
        self.gradsD = [FlatGrads(d.parameters(), o) for d, o in zip(netsD, optimizersD)]
        # data parallelism (the nn.DataParallel replacement, cub:139,163): one gradient reducer per network when the
        # process group has several ranks (data_parallel=None: decided by the process group; False: never).  A
        # discriminator's reducer starts from the tail of its flat buffer while its backward is still running.
        from . import parallel
        dp = parallel.world()[1] > 1 if data_parallel is None else data_parallel
        self.redG = parallel.make_reducer(self.gradsG.flat, netG, self.gradsG.params, self.gradsG.offsets) if dp else None
        self.redD = [parallel.make_reducer(g.flat, d, g.params, g.offsets) if dp else None for d, g in zip(netsD, self.gradsD)]
        # the discriminators' optimiser steps overlap their backward passes (EKL_TAIL_ADAM=0: one Adam launch after it)
        tail = os.environ.get("EKL_TAIL_ADAM", "1") != "0" and all(hasattr(o, "apply_slice") and o.flat_p.is_cuda for o in optimizersD)
        self.tailD = [TailUpdate(o, d, g, r) if tail else None for o, d, g, r in zip(optimizersD, netsD, self.gradsD, self.redD)]
        # the generator's tail: everything but its conditioning nets (and, one mark earlier, everything but the stem Linear) is
        # final before backward reaches them (model._GBase)
        self.tailG = TailUpdate(optimizerG, netG, self.gradsG, self.redG) if (
            tail and hasattr(optimizerG, "apply_slice") and optimizerG.flat_p.is_cuda) else None
        if self.tailG is not None:
            ops.GRAD_MARKS[id(netG)] = self.tailG.on_mark
            if hasattr(netG, "h_net1"):
                ops.GRAD_MARKS[id(netG.h_net1)] = self.tailG.on_mark
        for d, r, t in zip(netsD, self.redD, self.tailD):
            if t is not None:
                ops.GRAD_MARKS[id(d)] = t.on_mark
            elif r is not None:
                ops.GRAD_MARKS[id(d)] = r.on_mark
        self.bn_counters = BnCounters([netG] + list(netsD))
        # EKL_PARALLEL_D=0 runs the discriminators one after the other on the caller's stream
        self.parallel_d = os.environ.get("EKL_PARALLEL_D", "1") != "0" and len(netsD) > 1
        self.d_streams = None
        self._pack_sides = []               # (issuing stream, side stream) of operand refreshes still to be joined
        self.d_logits = {}                   # idx -> (real, wrong, fake) x [match p, uncond p, class log-probs]
        self.uncond = float(cfg.TRAIN.COEFF.UNCOND_LOSS)
        self.kl_coeff = float(cfg.TRAIN.COEFF.KL)
        self.color_coeff = float(cfg.TRAIN.COEFF.COLOR_LOSS)      # 0.0 in every shipped yml (config.py:61)
        self.last_color = []
        self.cat_z = cfg.TRAIN.CAT_Z

    @staticmethod
    def _apply(opt, red):
        """Optimiser step on the (rank-averaged) gradients: with a reducer, its remaining slices go out, the side stream
        is joined and the optimiser reads the buffer the reducer hands back (bf16 staging or the fp32 flat buffer)."""
        if red is None:
            opt.step()
            return
        g, is_bf16 = red.finish()
        if is_bf16:
            opt.step(grads_bf16=g)
        else:
            opt.step()

    def _d_begin(self, idx):
        """Arms the overlap machinery of discriminator idx right before its backward."""
        if self.tailD[idx] is not None:
            self.tailD[idx].begin()
        elif self.redD[idx] is not None:
            self.redD[idx].begin()

    def _d_apply(self, idx):
        """Optimiser step of discriminator idx, then the refresh of its packed filter operands on a side stream (the
        generator-loss pass through the updated discriminator follows on this branch; it reaches the operands that need
        packing -- data-gradient filters -- only in its backward)."""
        if self.tailD[idx] is not None:
            self.tailD[idx].finish()
        else:
            self._apply(self.optsD[idx], self.redD[idx])
        if next(self.netsD[idx].parameters()).is_cuda:
            side = ops.prepack(self.netsD[idx].parameters(), "D%d" % idx)
            if side is not None:
                self._pack_sides.append((torch.cuda.current_stream(), side))

    # ---- (1) generate: cub:567-587 / trainer.py:524-528
    def generate(self, noise, txt, cls_cond, eps=None, seed=None):
        dev = noise.device
        if dev.type == "cuda":
            ops.ARENA.reset(dev)             # one memset per step: zeroed scratch of every BatchNorm statistics buffer
            # the generator's filters changed at the end of the previous step: refresh their packed operands on a side
            # stream, behind the conditioning nets / stem that open the forward pass
            side = ops.prepack(self.netG.parameters(), "G")
        with tensor_core_matmul():
            out = self._generate(noise, txt, cls_cond, eps, seed)
        if dev.type == "cuda" and side is not None:
            torch.cuda.current_stream().wait_stream(side)
        return out

    def _generate(self, noise, txt, cls_cond, eps=None, seed=None):
        if self.kind == "catz_ca":
            if hasattr(self.netG, "vc_net1"):      # COND_G_NET_CATZ (model.py:567): both codes from VC_NETs, two seed draws
                out = self.netG(noise, txt, cls_cond, seed1=eps, seed2=seed)
            else:
                out = self.netG(noise, txt, cls_cond, eps=eps, seed=seed)
            (self.hcodes, self.mu1, self.mu2, self.logvar1, self.logvar2, self.std1, self.std2) = out
            if self.cat_z == "concat":
                self.mu = torch.cat((self.mu1, self.mu2), 1)
            elif self.cat_z == "product":
                self.mu = self.mu1 * self.mu2
            else:
                self.mu = self.mu1 + self.mu2
            self.kls = [(self.mu1, self.logvar1), (self.mu2, self.logvar2)]
        else:
            cond_info = torch.cat((txt, cls_cond), 1) if self.cond == "txt+cls" else txt
            self.hcodes, self.mu, self.logvar, self.std = self.netG(noise, cond_info, seed=seed)
            self.kls = [(self.mu, self.logvar)]
        self.fake_imgs = self.netG.image(self.hcodes)
        return self.fake_imgs

    # ---- (2) one discriminator update: cub:404-461
    def d_step(self, idx, real_imgs, wrong_imgs, real_cp, fake_cp):
        with tensor_core_matmul():
            out = self._d_step(idx, real_imgs, wrong_imgs, real_cp, fake_cp)
        self._join_packs()
        return out

    def _d_step(self, idx, real_imgs, wrong_imgs, real_cp, fake_cp):
        netD, opt, grads = self.netsD[idx], self.optsD[idx], self.gradsD[idx]
        B = real_imgs.shape[0]
        grads.zero()
        if hasattr(netD, "heads_raw") and self.uncond > 0:
            # fused path: raw logits of the stacked real / wrong / fake pass -> one loss kernel (cub:423-448)
            lm, lu, lc = netD.heads_raw((real_imgs, wrong_imgs, self.fake_imgs[idx].detach()), self.mu.detach(), groups=3)
            losses, pm, pu, logp = ops.d_loss(lm, lu, lc, real_cp, fake_cp, 3, B, (1, 0, 0), (1, 1, 0), (0, -1, 1), self.uncond)
            self._d_begin(idx)
            losses[0].backward()
            self._d_update(idx)
            self.d_logits[idx] = tuple([pm[i * B:(i + 1) * B], pu[i * B:(i + 1) * B], logp[i * B:(i + 1) * B]] for i in range(3))
            d = losses.detach()
            return d[0], d[1], d[2], d[3]
        # the three reference forwards (real / wrong / fake) as one pass; the batches are gathered by the stem kernel
        out = netD((real_imgs, wrong_imgs, self.fake_imgs[idx].detach()), self.mu.detach(), groups=3)
        real, wrong, fake = [[o[i * B:(i + 1) * B] for o in out] for i in range(3)]
        errD_match = _bce_const(real[0], 1) + _bce_const(wrong[0], 0) + _bce_const(fake[0], 0)
        if len(out) > 1 and self.uncond > 0:
            u = self.uncond
            errD_uncond = u * _bce_const(real[1], 1) + u * _bce_const(wrong[1], 1) + u * _bce_const(fake[1], 0)
            # two-head discriminators (D_NET64/128/256, model.py:874-914: [cond, uncond], no class head) have no loss
            # assembly in the reference (its train_joint_Dnet indexes a third output); for them the
            # same match + uncond sum is taken -- the StackGAN++ form the reference keeps as comments (trainer.py:408-410)
            errD_cls = (ce_loss(real[2], real_cp) + ce_loss(fake[2], fake_cp)) if len(out) > 2 else torch.zeros((), device=real_imgs.device)
            errD = errD_match + errD_uncond + errD_cls
        else:
            errD_uncond = errD_cls = torch.zeros((), device=real_imgs.device)
            errD = _bce_const(real[0], 1) + 0.5 * (_bce_const(wrong[0], 0) + _bce_const(fake[0], 0))
        self._d_begin(idx)
        errD.backward()
        self._d_update(idx)
        self.d_logits[idx] = (real, wrong, fake)
        return errD, errD_match, errD_uncond, errD_cls

    @property
    def last_d_logits(self):
        """Logits of the deepest discriminator's latest update (real / wrong / fake groups)."""
        return self.d_logits[max(self.d_logits)]

    def _d_update(self, idx):
        """Gradient average over the ranks (N > 1: the slices still outstanding, see parallel.GradReducer) + Adam step of
        discriminator idx.  The updates of the discriminators are independent (cub:594-596 loops over them) and run as
        parallel stream branches (d_steps), so one discriminator's exchange also overlaps the others' compute."""
        self._d_apply(idx)

    def _join_packs(self):
        """Side-stream operand refreshes rejoin the stream they were forked from (a captured step must end with every
        stream joined; consumers already ordered themselves individually)."""
        for main, side in self._pack_sides:
            main.wait_stream(side)
        self._pack_sides = []

    def _streams(self):
        if self.d_streams is None:
            # the deepest discriminator's branch is the critical path of the parallel section: a high-priority stream lets
            # its kernels win SMs over the fillers (captured kernel nodes inherit the stream priority; measured on config
            # 2: 8.56 -> 8.42 ms/step)
            last = len(self.netsD) - 1
            self.d_streams = [torch.cuda.Stream(priority=-1 if i == last else 0) for i in range(len(self.netsD))]
        return self.d_streams

    def d_steps(self, real_imgs, wrong_imgs, real_cp, fake_cp):
        """All discriminator updates of one iteration (cub:594-596).  They are independent of each other (own
        parameters, gradients, optimiser; inputs are the detached fakes), so each runs on its own stream: the
        latency-bound tails of one discriminator (4x4 / 8x8 maps: kernels of 50-150 CTAs) fill the SMs another one
        leaves idle, and with several ranks its gradient exchange hides behind the others' compute.  Captured into the
        step graph as parallel branches.  Returns the per-discriminator loss tuples in index order."""
        n = len(self.netsD)
        errDs = [None] * n
        if not self.parallel_d:
            for i in reversed(range(n)):
                errDs[i] = self.d_step(i, real_imgs[i], wrong_imgs[i], real_cp, fake_cp)
            return errDs
        main = torch.cuda.current_stream()
        streams = self._streams()
        for i in reversed(range(n)):              # largest first
            streams[i].wait_stream(main)
            with torch.cuda.stream(streams[i]):
                errDs[i] = self.d_step(i, real_imgs[i], wrong_imgs[i], real_cp, fake_cp)
        for st in streams:
            main.wait_stream(st)
        return errDs

    def update(self, real_imgs, wrong_imgs, real_cp, fake_cp):
        """Everything after generate(): the discriminator updates (cub:594-596) and the generator update through the
        updated discriminators (cub:604-608).  With parallel branches each discriminator's stream runs its update AND,
        right behind it, the generator-loss forward through that same (now updated) discriminator -- the small
        discriminators' generator-loss passes fill the SMs while the largest one is still updating, instead of all
        three waiting for a common join.  Returns (errDs, errG)."""
        n = len(self.netsD)
        fused = self.parallel_d and self.uncond > 0 and all(hasattr(d, "heads_raw") for d in self.netsD)
        if not fused:
            return self.d_steps(real_imgs, wrong_imgs, real_cp, fake_cp), self.g_step(real_cp)
        errDs, per_d = [None] * n, [None] * n
        main = torch.cuda.current_stream()
        streams = self._streams()
        with tensor_core_matmul():
            try:
                for i in reversed(range(n)):              # largest first
                    streams[i].wait_stream(main)
                    with torch.cuda.stream(streams[i]):
                        errDs[i] = self._d_step(i, real_imgs[i], wrong_imgs[i], real_cp, fake_cp)
                        netD = self.netsD[i]
                        netD.requires_grad_(False)        # the reference computes and discards these
                        lm, lu, lc = netD.heads_raw(self.fake_imgs[i], self.mu)
                        per_d[i] = ops.d_loss(lm, lu, lc, real_cp, None, 1, lm.shape[0], (1,), (1,), (0,), self.uncond)
                self._join_packs()
                for st in streams:
                    main.wait_stream(st)
                errG = self._g_update(real_cp, per_d)
            finally:
                for d in self.netsD:
                    d.requires_grad_(True)
        return errDs, errG

    def color_consistency(self, fake_imgs):
        """Colour-consistency regulariser between adjacent stages, the StackGAN++ term the reference keeps the statistics
        function and the coefficient for (cub:33-52 compute_mean_covariance, config.py:61 COEFF.COLOR_LOSS) but never
        assembles:  sum_{i>=1} coeff * ( MSE(mu_i, mu_{i-1}) + 5 * MSE(cov_i, cov_{i-1}) ), the lower stage detached.
        Each stage's statistics are computed once (ekl_color_stats_fwd); the lower role uses them detached."""
        stats = []
        for i, img in enumerate(fake_imgs):
            if i == 0:
                with torch.no_grad():
                    stats.append(compute_mean_covariance(img))
            else:
                stats.append(compute_mean_covariance(img))
        total, self.last_color = 0, []
        for i in range(1, len(fake_imgs)):
            (mu1, cov1), (mu2, cov2) = stats[i], stats[i - 1]
            like_mu = self.color_coeff * F.mse_loss(mu1, mu2.detach())
            like_cov = self.color_coeff * 5 * F.mse_loss(cov1, cov2.detach())
            self.last_color.append((like_mu.detach(), like_cov.detach()))
            total = total + like_mu + like_cov
        return total

    # ---- (3) generator loss through the UPDATED discriminators: cub:463-490
    def g_loss(self, real_cp, per_d=None):
        """per_d: the discriminators' fused loss tuples when their forward passes already ran (update())."""
        errGs_match = errGs_uncond = errGs_cls = errGs_total_fused = 0
        self.last_g_logits = []
        main = torch.cuda.current_stream()
        par = self.parallel_d and all(hasattr(d, "heads_raw") for d in self.netsD) and self.uncond > 0
        done = per_d is not None
        if done:
            par = True
        else:
            per_d = []
        if par and not done:
            # the discriminators judge their own stage's fake independently: one stream each (autograd replays every
            # branch's backward on the stream of its forward), joined before the losses are summed
            for i, netD in enumerate(self.netsD):
                st = self._streams()[i]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    lm, lu, lc = netD.heads_raw(self.fake_imgs[i], self.mu)
                    per_d.append(ops.d_loss(lm, lu, lc, real_cp, None, 1, lm.shape[0], (1,), (1,), (0,), self.uncond))
            for st in self._streams():
                main.wait_stream(st)
        for i, netD in enumerate(self.netsD):
            if hasattr(netD, "heads_raw") and self.uncond > 0:
                if par:
                    losses, pm, pu, logp = per_d[i]
                else:
                    lm, lu, lc = netD.heads_raw(self.fake_imgs[i], self.mu)
                    losses, pm, pu, logp = ops.d_loss(lm, lu, lc, real_cp, None, 1, lm.shape[0], (1,), (1,), (0,), self.uncond)
                d = losses.detach()
                errGs_match, errGs_uncond, errGs_cls = errGs_match + d[1], errGs_uncond + d[2], errGs_cls + d[3]
                errGs_total_fused = errGs_total_fused + losses[0]
                self.last_g_logits.append([pm, pu, logp])
                continue
            outputs = netD(self.fake_imgs[i], self.mu)
            errGs_total_fused = errGs_total_fused + _bce_const(outputs[0], 1)
            errGs_match = errGs_match + _bce_const(outputs[0], 1)
            if len(outputs) > 1 and self.uncond > 0:
                u_ = self.uncond * _bce_const(outputs[1], 1)
                c_ = ce_loss(outputs[2], real_cp) if len(outputs) > 2 else torch.zeros((), device=u_.device)
                errGs_uncond, errGs_cls = errGs_uncond + u_, errGs_cls + c_
                errGs_total_fused = errGs_total_fused + u_ + c_
            self.last_g_logits.append(outputs)
        kl = [self._kl(m, lv) for m, lv in self.kls]
        # only losses[0] of the fused kernel carries gradient: the total is assembled from it (components are reported)
        errG_total = errGs_total_fused + sum(kl) * self.kl_coeff
        if self.color_coeff > 0 and len(self.fake_imgs) > 1:
            errG_total = errG_total + self.color_consistency(self.fake_imgs)
        return (errG_total, errGs_match, errGs_uncond, errGs_cls) + tuple(kl)

    def _kl(self, mu, logvar):
        """KL term of one conditioning net: the value its fused reparameterisation kernel already produced for exactly
        these (mu, logvar) tensors, else cub:54-58 in torch."""
        for m in self.netG.modules():
            of = getattr(m, "last_kl_of", None)
            if of is not None and of[0] is mu and of[1] is logvar:
                return m.last_kl
        return KL_loss(mu, logvar)

    def g_step(self, real_cp):
        with tensor_core_matmul():
            return self._g_step(real_cp)

    def _g_step(self, real_cp):
        for d in self.netsD:
            d.requires_grad_(False)          # the reference computes and discards these
        try:
            return self._g_update(real_cp)
        finally:
            for d in self.netsD:
                d.requires_grad_(True)

    def _g_update(self, real_cp, per_d=None):
        self.gradsG.zero()
        res = self.g_loss(real_cp, per_d)
        if self.tailG is not None:
            self.tailG.begin()
        res[0].backward()
        self._join_packs()
        if self.tailG is not None:
            self.tailG.finish()
        else:
            if self.redG is not None:
                self.redG.begin()
            self._apply(self.optG, self.redG)
        self.bn_counters.flush()
        return res

    # ---- whole step on device-resident inputs
    def step(self, real_imgs, wrong_imgs, txt, cls_cond, real_cp, fake_cp, noise, eps=None, seed=None):
        self.generate(noise, txt, cls_cond, eps, seed)
        # the discriminator updates are independent (cub:594-596): parallel stream branches, the largest issued first
        return self.update(real_imgs, wrong_imgs, real_cp, fake_cp)


class GraphedStep:
    """The whole training step captured once into a CUDA graph (generate -> D updates -> G update, Adam and the
    gradient all-reduce included) and replayed per batch: ~10^3 kernel launches per step become one graph launch.

    step(data): copies the loader's host batch into static device buffers (async from pinned memory), replays, and
    returns the static loss tensors (errDs [num_Ds,4], errG [4+]).  RNG (noise, CA eps, VC seed) is drawn on the
    device inside the graph."""

    def __init__(self, trainer, example_data, warmup=3):
        self.tr = trainer
        dev = trainer.device
        imgs, wrong, emb, cls, _ = example_data
        n = trainer.num_Ds
        # the graph's input buffers are slices of ONE allocation, mirrored by a staging allocation: a prefetched batch
        # moves from staging into place with a single device copy (see prefetch)
        hosts = [imgs[i] for i in range(n)] + [wrong[i] for i in range(n)] + [emb, cls]
        offs, total = [], 0
        for t in hosts:
            offs.append(total)
            total += (t.numel() * t.element_size() + 255) // 256 * 256
        self._in_flat = torch.empty(total, dtype=torch.uint8, device=dev)
        self._stage_flat = torch.empty(total, dtype=torch.uint8, device=dev)

        def carve(flat):
            return [flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for o, t in zip(offs, hosts)]
        ins, self._stage = carve(self._in_flat), carve(self._stage_flat)
        self.s_imgs, self.s_wrong, self.s_emb, self.s_cls = ins[:n], ins[n:2 * n], ins[2 * n], ins[2 * n + 1]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in hosts)
        self.copy_stream = torch.cuda.Stream()
        self._staged, self._ev_staged, self._ev_consumed = None, torch.cuda.Event(), torch.cuda.Event()
        self.load(example_data)
        B, dv = emb.shape[0], dev
        self.eps = torch.zeros(B, cfg.GAN.EMBEDDING_DIM, device=dv)
        self.seed = torch.zeros(B, cfg.GAN.MANIFD_DIM, device=dv)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES[0]
        # thread_local: a DataLoader's pin-memory thread may allocate page-locked memory while the capture is open
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = self._body()
        self.launches_per_step = ops.LAUNCHES[0] - n0
        torch.cuda.synchronize()

    def _body(self):
        tr = self.tr
        tr.real_imgs, tr.wrong_imgs, tr.txt_embedding = self.s_imgs, self.s_wrong, self.s_emb
        tr.cls_label = (self.s_cls.long() - 1) if getattr(tr, "CLS_KIND", "index") == "index" else self.s_cls
        tr.noise.normal_(0, 1)
        self.eps.normal_(0, 1)
        self.seed.normal_(0, 1)
        tr.generate(self.eps, self.seed)
        errDs, errG = tr.engine.update(tr.real_imgs, tr.wrong_imgs, tr.real_cp, tr.fake_cp)
        return torch.stack([torch.stack([x.detach().float() for x in e]) for e in errDs]), \
            torch.stack([x.detach().float() for x in errG])

    def load(self, data):
        imgs, wrong, emb, cls, _ = data
        for i in range(self.tr.num_Ds):
            self.s_imgs[i].copy_(imgs[i], non_blocking=True)
            self.s_wrong[i].copy_(wrong[i], non_blocking=True)
        self.s_emb.copy_(emb, non_blocking=True)
        self.s_cls.copy_(cls, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.out

    def accepts(self, data):
        """True when the host batch has the shapes / dtypes the graph was captured for."""
        imgs, wrong, emb, cls, _ = data
        n = self.tr.num_Ds
        srcs = [imgs[i] for i in range(n)] + [wrong[i] for i in range(n)] + [emb, cls]
        return all(s.shape == d.shape and s.dtype == d.dtype for s, d in zip(srcs, self._stage))

    def prefetch(self, data):
        """Start the host->device copy of a FUTURE batch on the copy stream into the staging buffers; it overlaps the
        step that is running.  The step(data) call that later receives this same batch object moves it into the graph's
        input buffers with one device-to-device copy (~50 MB: tens of microseconds) instead of a PCIe transfer on the
        critical path.  The host tensors must stay unchanged until then (a DataLoader batch does)."""
        imgs, wrong, emb, cls, _ = data
        n = self.tr.num_Ds
        cs = self.copy_stream
        cs.wait_event(self._ev_consumed)          # the previously staged batch has left the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage, [imgs[i] for i in range(n)] + [wrong[i] for i in range(n)] + [emb, cls]):
                dst.copy_(src, non_blocking=True)
            self._ev_staged.record(cs)
        self._staged = data

    def step(self, data, next_data=None):
        """One training step on the loader's host batch.  next_data (optional): the batch of the following call; its
        upload is started behind this step's graph launch so that it never waits on PCIe."""
        if self._staged is data:
            main = torch.cuda.current_stream()
            main.wait_event(self._ev_staged)
            self._in_flat.copy_(self._stage_flat, non_blocking=True)
            self._ev_consumed.record(main)
            self._staged = None
        else:
            self.load(data)
        out = self.replay()
        if next_data is not None:
            self.prefetch(next_data)
        return out

    def run(self, loader):
        """Iterate a loader with the upload of batch k+1 hidden behind step k; yields (errDs, errG) per batch."""
        it = iter(loader)
        cur = next(it, None)
        while cur is not None:
            nxt = next(it, None)
            yield self.step(cur, nxt)
            cur = nxt
