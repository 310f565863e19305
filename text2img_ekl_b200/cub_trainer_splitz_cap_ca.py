"""Drop-in for the reference's cub_trainer_splitz_cap_ca.py (split-z + capsule + CA_NET flavour): same module
functions (`load_network`, `define_optimizers`, `KL_loss`, `ce_loss`, `compute_mean_covariance`, `weights_init`) and
the same `condGANTrainer` methods (`prepare_data`, `onehot`, `train_joint_Dnet`, `loss_joint_Gnet`, `train`), driving
the B200 kernels through engine.StepEngine.  One process per GPU: `nn.DataParallel` (cub:139,163) is replaced by
torch.distributed (NCCL) gradient all-reduce when WORLD_SIZE > 1; BatchNorm statistics stay per replica.
Out of scope: Inception-score evaluation, tensorboard image summaries.
"""
import os
import time

import torch
import torch.nn as nn
import torch.optim as optim

from . import model
from . import ops
from .engine import (KL_loss, StepEngine, ce_loss, compute_mean_covariance, onehot)  # noqa: F401  (reference names)
from .datasets import stage_images
from .miscc.config import cfg
from .miscc.losslog import AsyncLossLog
from .miscc.utils import mkdir_p
from . import parallel

USE_CLS = True
SPLIT_Z = True


def weights_init(m):
    """cub:67-77: orthogonal(gain 1) for Conv / Linear class names (incl. CapsuleLinear), BN gamma ~ N(1, 0.02), beta 0."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1 and hasattr(m, "weight"):
        nn.init.orthogonal_(m.weight.data, 1.0)
    elif classname.find("BatchNorm") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif classname.find("Linear") != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def copy_G_params(net):
    return [p.data.clone() for p in net.parameters()]


def load_params(net, new_param):
    for p, new_p in zip(net.parameters(), new_param):
        p.data.copy_(new_p)


def build_G(use_cap=None, g_class=None):
    """The generator train() builds (cub:130-135); evaluate() passes cfg.TEST.G_CAPSULE (cub:787).  g_class names
    another split-z assembly of model.py to drive through the same step (COND_G_NET_CATZ, model.py:567)."""
    use_cap = cfg.TRAIN.G_CAPSULE if use_cap is None else use_cap
    shareGs = model.get_shareGs(cfg.GAN.GF_DIM)
    if g_class is not None:
        netG = getattr(model, g_class)(cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM, shareGs, use_cap=use_cap,
                                       cat=cfg.TRAIN.CAT_Z, exchange=cfg.TRAIN.EXCHANGE)
    elif USE_CLS and SPLIT_Z:
        netG = model.COND_G_NET_CATZ_CA(cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM, shareGs, use_cap=use_cap,
                                        cat=cfg.TRAIN.CAT_Z, exchange=cfg.TRAIN.EXCHANGE)            # cub:130
    else:
        netG = model.COND_G_NET(cfg.TEXT.DIMENSION, shareGs, use_cap=use_cap)                       # cub:135
    return netG, shareGs


def _strip_module(state_dict):
    """Reference snapshots are saved from the nn.DataParallel wrapper: keys carry a 'module.' prefix (cub:663)."""
    return {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}


def save_model(netG, avg_param_G, netsD, epoch, model_dir):
    """cub:218-228: netG_<epoch>.pth (with avg_param_G loaded first when given) and netD<i>.pth, the files
    cfg.TRAIN.NET_G / cfg.TRAIN.NET_D point load_network at.  Keys carry the reference's 'module.' prefix."""
    if avg_param_G is not None:
        load_params(netG, avg_param_G)
    wrap = lambda net: {"module." + k: v.detach().cpu() for k, v in net.state_dict().items()}
    torch.save(wrap(netG), "%s/netG_%d.pth" % (model_dir, epoch))
    for i, netD in enumerate(netsD):
        torch.save(wrap(netD), "%s/netD%d.pth" % (model_dir, i))


def to_uint8_nhwc(images):
    """cub:771-772: [-1, 1] float NCHW -> uint8 NHWC (add 1, /2, *255, clamp, truncate), computed where the images live."""
    return images.detach().add(1).div(2).mul(255).clamp(0, 255).byte().permute(0, 2, 3, 1).contiguous()


def image_grid_uint8(images, nrow=10, padding=2):
    """The tile sheet the reference writes through torchvision's save_image(nrow=10, normalize=True) (cub:756): min-max
    normalisation over the whole stack, `nrow` tiles per row, 2-pixel black gutters, round to uint8.  [N,3,H,W] -> HWC."""
    x = images.detach().float()
    lo, hi = x.min(), x.max()
    x = (x - lo) / (hi - lo).clamp_min(1e-5)
    n, c, h, w = x.shape
    cols = min(nrow, n)
    rows = (n + cols - 1) // cols
    sheet = x.new_zeros(c, rows * (h + padding) + padding, cols * (w + padding) + padding)
    for k in range(n):
        r, q = divmod(k, cols)
        sheet[:, padding + r * (h + padding):padding + r * (h + padding) + h,
              padding + q * (w + padding):padding + q * (w + padding) + w] = x[k]
    return sheet.mul(255).add(0.5).clamp(0, 255).byte().permute(1, 2, 0).contiguous()


def build_Ds(plain=False):
    """cub:141-157 (JOINT_D_NET*).  plain: the StackGAN++ two-head D_NET64/128/256 (model.py:874-914, 1006-1050,
    1154-1202) -- in scope as modules; no shipped reference trainer builds them."""
    if plain:
        return [D() for D in (model.D_NET64, model.D_NET128, model.D_NET256)[:cfg.TREE.BRANCH_NUM]]
    netsD = []
    if cfg.TREE.BRANCH_NUM > 0:
        netsD.append(model.JOINT_D_NET64(use_cap=cfg.TRAIN.D_CAPSULE))
    if cfg.TREE.BRANCH_NUM > 1:
        netsD.append(model.JOINT_D_NET128(use_cap=cfg.TRAIN.D_CAPSULE) if cfg.TREE.SCALE == 2 else model.JOINT_D_NET256())
    if cfg.TREE.BRANCH_NUM > 2:
        # the reference asserts 'br3 todo' here (cub:156); the 3-stage config composes JOINT_D_NET256 (cfg 2)
        netsD.append(model.JOINT_D_NET256())
    return netsD


def load_snapshots(netG, netsD):
    """cub:171-184 / trainer.py:138-152: load cfg.TRAIN.NET_G into the generator and cfg.TRAIN.NET_D<i>.pth into the
    discriminators; the start count is the number in the generator file name + 1 (the reference's int() of the name
    tail fails on the 'netG_epoch%d.pth' files its own loop writes, SURVEY app. A #11: only the digits are used here)."""
    count = 0
    if cfg.TRAIN.NET_G != "":
        state_dict = torch.load(cfg.TRAIN.NET_G, map_location="cpu")
        netG.load_state_dict(_strip_module(state_dict))
        name = os.path.basename(cfg.TRAIN.NET_G)
        digits = "".join(ch for ch in name[name.rfind("_") + 1:name.rfind(".")] if ch.isdigit())
        count = int(digits) + 1 if digits else 0
    if cfg.TRAIN.NET_D != "":
        for i, d in enumerate(netsD):
            sd = torch.load("%s%d.pth" % (cfg.TRAIN.NET_D, i), map_location="cpu")
            d.load_state_dict(_strip_module(sd))
    return count


def load_network(gpus, device=None, g_class=None, plain_d=False):
    """cub:113-196.  Returns (netG, shareGs, netsD, num_Ds, count)."""
    device = device or (torch.device("cuda", gpus[0]) if gpus else torch.device("cuda"))
    netG, shareGs = build_G(g_class=g_class)
    netG.apply(weights_init)
    netsD = build_Ds(plain_d)
    for d in netsD:
        d.apply(weights_init)
    count = load_snapshots(netG, netsD)
    netG.to(device)
    model.to_kernel_layout(netG)
    for d in netsD:
        d.to(device)
        model.to_kernel_layout(d)
    return netG, shareGs, netsD, len(netsD), count


def define_optimizers(netG, netsD=()):
    """cub:199-215: Adam(lr 2e-4, betas (0.5, 0.999)) per network.  On the GPU: optim.FlatAdam, the same update as one
    fused kernel per network over flat parameter / gradient / moment buffers (graph-capturable)."""
    kw = dict(betas=(0.5, 0.999))
    if next(netG.parameters()).is_cuda:
        from .optim import FlatAdam
        optimizersD = [FlatAdam(d.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, **kw) for d in netsD]
        optimizerG = FlatAdam(netG.parameters(), lr=cfg.TRAIN.GENERATOR_LR, **kw)
        return optimizerG, optimizersD
    optimizersD = [optim.Adam(d.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, **kw) for d in netsD]
    optimizerG = optim.Adam(netG.parameters(), lr=cfg.TRAIN.GENERATOR_LR, **kw)
    return optimizerG, optimizersD


class condGANTrainer(object):
    KIND = "catz_ca"
    COND = "txt+cls"
    G_CLASS = None                 # None: the generator cub:130-135 builds
    PLAIN_D = False                # True: the two-head D_NET* instead of JOINT_D_NET*

    def __init__(self, output_dir, data_loader, imsize):
        if cfg.TRAIN.FLAG and output_dir:
            self.model_dir = os.path.join(output_dir, "Model")
            self.image_dir = os.path.join(output_dir, "Image")
            self.log_dir = os.path.join(output_dir, "Log")
            for d in (self.model_dir, self.image_dir, self.log_dir):
                mkdir_p(d)
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        s_gpus = str(cfg.GPU_ID).split(",")
        self.gpus = [local_rank] if "LOCAL_RANK" in os.environ else [int(ix) for ix in s_gpus]
        self.num_gpus = len(self.gpus)
        self.device = torch.device("cuda", self.gpus[0])
        torch.cuda.set_device(self.device)
        self.batch_size = cfg.TRAIN.BATCH_SIZE
        self.max_epoch = cfg.TRAIN.MAX_EPOCH
        self.snapshot_interval = cfg.TRAIN.SNAPSHOT_INTERVAL
        self.data_loader = data_loader
        self.num_batches = len(data_loader) if data_loader is not None else 0
        self.engine = None

    # ------------------------------------------------------------------ data
    def prepare_data(self, data):
        """cub:295-320: zero-base the class index and move the batch to the device."""
        imgs, w_imgs, t_embedding, cls, _ = data
        cls = cls.long() - 1
        dev = self.device
        real_vimgs = stage_images(imgs, self.num_Ds, dev)       # fp32 pyramid as delivered, or uint8 crops -> device pyramid
        wrong_vimgs = stage_images(w_imgs, self.num_Ds, dev)
        return imgs, real_vimgs, wrong_vimgs, t_embedding.to(dev, non_blocking=True), cls.to(dev, non_blocking=True)

    def onehot(self, cls_vec, n):
        return onehot(cls_vec, n)

    # ------------------------------------------------------------------ setup (cub:494-537)
    def _replicate(self):
        """One process per GPU (the nn.DataParallel replacement, cub:139,163): join the process group torchrun described
        in the environment, start every replica from rank 0's weights and buffers (DataParallel replicates GPU0's), and
        give each rank its own device RNG stream for the in-step noise / eps / seed draws.  No-op for one process."""
        rank, ws = parallel.init_from_env()
        if ws > 1:
            parallel.broadcast_params([self.netG] + list(self.netsD))
            if self.device.type == "cuda":
                torch.cuda.manual_seed(torch.cuda.initial_seed() + rank)
            else:
                torch.manual_seed(torch.initial_seed() + rank)
        self.rank, self.world_size = rank, ws

    def setup(self):
        self.netG, self.shareGs, self.netsD, self.num_Ds, start_count = load_network(self.gpus, self.device, self.G_CLASS, self.PLAIN_D)
        self._replicate()                    # before the optimisers re-home the parameters into flat buffers
        self.optimizerG, self.optimizersD = define_optimizers(self.netG, self.netsD)
        self.criterion = nn.BCELoss()
        self.CE = ce_loss
        B = self.batch_size
        self.real_labels = torch.ones(B, device=self.device)
        self.fake_labels = torch.zeros(B, device=self.device)
        self.fake_cp = torch.zeros(B, cfg.GAN.ENTITY_DIM + 1, device=self.device)
        self.fake_cp[:, -1] = 1
        self.noise = torch.zeros(B, cfg.GAN.Z_DIM, device=self.device)
        self.engine = StepEngine(self.netG, self.netsD, self.optimizerG, self.optimizersD, self.KIND, self.COND)
        return start_count

    # ------------------------------------------------------------------ reference-named step pieces
    def labels(self):
        """cub:556-557 (birds): one-hot class condition [B,E] and class target [B,E+1]."""
        self.cls_onehot = self.onehot(self.cls_label, cfg.GAN.ENTITY_DIM)
        self.real_cp = self.onehot(self.cls_label, cfg.GAN.ENTITY_DIM + 1)
        return self.cls_onehot

    def generate(self, eps=None, seed=None):
        e = self.engine
        self.fake_imgs = e.generate(self.noise, self.txt_embedding, self.labels(), eps, seed)
        self.hcodes, self.mu = e.hcodes, e.mu
        for k in ("mu1", "mu2", "logvar1", "logvar2", "std1", "std2", "logvar", "std"):
            if hasattr(e, k):
                setattr(self, k, getattr(e, k))

    def train_joint_Dnet(self, idx, count):
        """cub:404-461 -> (errD, errD_match, errD_uncond, errD_cls)."""
        return self.engine.d_step(idx, self.real_imgs[idx], self.wrong_imgs[idx], self.real_cp, self.fake_cp)

    def loss_joint_Gnet(self, count):
        """cub:463-490 -> (errG_total, match, uncond, cls, kl_sen, kl_cls)."""
        return self.engine.g_loss(self.real_cp)

    def train_step(self, data, count=1, noise=None, eps=None, seed=None):
        """One iteration of the hot loop (cub:552-608)."""
        self.imgs_tcpu, self.real_imgs, self.wrong_imgs, self.txt_embedding, self.cls_label = self.prepare_data(data)
        if noise is None:
            self.noise.normal_(0, 1)
        else:
            self.noise.copy_(noise)
        self.generate(eps, seed)
        # the discriminator updates are independent: engine.update runs each as a stream branch that continues into the
        # generator-loss pass through that discriminator, then the generator update
        return self.engine.update(self.real_imgs, self.wrong_imgs, self.real_cp, self.fake_cp)

    EAGER_STEPS_BEFORE_CAPTURE = 2

    def _loop_step(self, data, nxt, count):
        """One iteration of train()'s loop.  On a GPU the first EAGER_STEPS_BEFORE_CAPTURE batches run eagerly (they
        initialise the library handles and buffers that must exist before a capture), then the whole step is captured
        once (engine.GraphedStep; the capture itself executes nothing) and every following full-size batch is one graph
        replay with the next batch's upload hidden behind it.  Batches of another shape (a loader without drop_last;
        the reference uses drop_last=True, main.py:135) take the eager path.  EKL_GRAPH=0 keeps everything eager."""
        if self.device.type == "cuda" and os.environ.get("EKL_GRAPH", "1") != "0":
            gs = getattr(self, "_graphed", None)
            if gs is None and getattr(self, "_eager_done", 0) >= self.EAGER_STEPS_BEFORE_CAPTURE:
                from .engine import GraphedStep
                gs = self._graphed = GraphedStep(self, data, warmup=0)
            if gs is not None and gs.accepts(data):
                return gs.step(data, nxt if nxt is not None and gs.accepts(nxt) else None)
        self._eager_done = getattr(self, "_eager_done", 0) + 1
        return self.train_step(data, count)

    def train(self):
        """cub:492-672.  On a GPU the loop runs on a side stream: autograd pins each parameter's gradient accumulation to
        the stream of its first use, and work pinned to the legacy default stream cannot be captured into a graph."""
        if getattr(self, "device", None) is None or torch.device(self.device).type != "cuda":
            return self._train_loop()
        cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._train_loop()
        cur.wait_stream(side)

    LOG_EVERY = 100                 # cub:457-460: D_loss scalars every 100 iterations

    def _log_scalars(self, count, errDs, errG):
        """The reference's per-100-iteration scalars (cub:457-460) without its `.item()` host syncs: an asynchronous
        device-to-host copy now, consumed when it has landed (miscc.losslog.AsyncLossLog)."""
        log = getattr(self, "loss_log", None)
        if log is None:
            path = os.path.join(self.log_dir, "scalars.jsonl") if hasattr(self, "log_dir") else None
            log = self.loss_log = AsyncLossLog(path, every=self.LOG_EVERY)
        if not log.due(count):
            log.poll()
            return
        d = errDs if torch.is_tensor(errDs) else torch.stack([torch.stack([x.detach().float() for x in e]) for e in errDs])
        g = errG if torch.is_tensor(errG) else torch.stack([x.detach().float() for x in errG])
        names = ["D_loss%d" % i for i in range(d.shape[0])] + ["G_loss"]
        log.push(count, names, torch.cat((d[:, 0].reshape(-1), g[:1].reshape(-1))))

    def _train_loop(self):
        start_count = self.setup()
        count = start_count
        start_epoch = start_count // max(self.num_batches, 1)
        for epoch in range(start_epoch, self.max_epoch):
            start_t = time.time()
            errDs = errG = None
            it = iter(self.data_loader)
            data = next(it, None)
            while data is not None:
                nxt = next(it, None)
                errDs, errG = self._loop_step(data, nxt, count)
                self._log_scalars(count, errDs, errG)
                count += 1
                data = nxt
            if errG is None:
                break
            end_t = time.time()
            self.loss_log.flush()
            tot = [sum(e[k] for e in errDs).item() for k in range(4)]
            print("[%d/%d][BN=%d][%d stages] Loss_D_all: %.2f match: %.2f uncond: %.2f cls: %.2f | "
                  "Loss_G_all: %.2f match: %.2f uncond: %.2f cls: %.2f KL: %s Time: %.2fs"
                  % (epoch, self.max_epoch, self.num_batches, self.num_Ds, tot[0], tot[1], tot[2], tot[3],
                     errG[0].item(), float(errG[1]), float(errG[2]), float(errG[3]),
                     " ".join("%.3f" % float(k) for k in errG[4:]), end_t - start_t))
            si = self.snapshot_interval
            if hasattr(self, "model_dir") and (epoch % si == si - 1 or epoch > 199):          # cub:664-669
                # keys carry the DataParallel "module." prefix like the reference's snapshots (cub:662-667); replicas
                # are identical, so rank 0 alone writes and the others wait for the file to be complete
                if getattr(self, "rank", 0) == 0:
                    sd = {"module." + k: v for k, v in self.netG.state_dict().items()}
                    torch.save(sd, "%s/netG_epoch%d.pth" % (self.model_dir, epoch))
                if getattr(self, "world_size", 1) > 1:
                    torch.distributed.barrier()

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, path, count):
        """Everything a bit-faithful resume needs in one file: G, every D (reference key layout, 'module.' prefix),
        the Adam state of each optimiser in torch.optim.Adam's state_dict format, and the iteration count.  The
        reference itself only snapshots netG (cub:662-669) and re-creates the optimisers on restart."""
        wrap = lambda net: {"module." + k: v.detach().cpu() for k, v in net.state_dict().items()}

        def opt_state(opt):         # compact host copies: the fused optimiser's moments are views into flat buffers
            sd = opt.state_dict()
            host = lambda v: v.detach().cpu().clone() if torch.is_tensor(v) else v
            return {"state": {i: {k: host(v) for k, v in st.items()} for i, st in sd["state"].items()},
                    "param_groups": sd["param_groups"]}
        torch.save({"netG": wrap(self.netG), "netsD": [wrap(d) for d in self.netsD],
                    "optimizerG": opt_state(self.optimizerG),
                    "optimizersD": [opt_state(o) for o in self.optimizersD], "count": int(count)}, path)

    def load_checkpoint(self, path):
        """Inverse of save_checkpoint, IN PLACE (parameter / moment storage keeps its addresses, so a captured step graph
        stays valid).  Returns the stored iteration count."""
        ck = torch.load(path, map_location="cpu")
        self.netG.load_state_dict(_strip_module(ck["netG"]))
        for d, sd in zip(self.netsD, ck["netsD"]):
            d.load_state_dict(_strip_module(sd))
        self.optimizerG.load_state_dict(ck["optimizerG"])
        for o, sd in zip(self.optimizersD, ck["optimizersD"]):
            o.load_state_dict(sd)
        for o in [self.optimizerG] + list(self.optimizersD):
            if hasattr(o, "refresh_shadow"):
                o.refresh_shadow()
        # packed filter operands that are not the optimiser's bf16 shadow (transposed data-gradient operands, up-conv tap
        # sums, filter windows) are refreshed by the step itself only AFTER each network's optimiser update: repack them
        # now, so that the first (possibly graph-replayed) step after a restore reads the restored filters everywhere
        if next(self.netG.parameters()).is_cuda:
            for net in [self.netG] + list(self.netsD):
                ops.mark_dirty(net.parameters())
                side = ops.prepack(net.parameters(), "restore")
                if side is not None:
                    torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        return int(ck["count"])

    # ------------------------------------------------------------------ generation path (cub:720-911)
    def save_singleimages(self, images, filenames, save_dir, split_dir, sentenceID, cls, imsize, noiseID):
        """cub:759-775: one PNG per sample, <save_dir>/single_samples/<key>_<size>_class<c>_sid<s>_nid<n>.png."""
        arr = to_uint8_nhwc(images).cpu().numpy()
        from PIL import Image
        for i in range(arr.shape[0]):
            stem = "%s/single_samples/%s" % (save_dir, filenames[i])
            mkdir_p(os.path.dirname(stem))
            Image.fromarray(arr[i]).save("%s_%d_class%d_sid%d_nid%d.png" % (stem, imsize, int(cls[i]), sentenceID, noiseID))

    def save_superimages(self, images_list, filenames, save_dir, split_dir, imsize):
        """cub:733-756: per sample, one sheet of its images for every sentence (10 per row, min-max normalised)."""
        from PIL import Image
        for i in range(images_list[0].size(0)):
            stem = "%s/super/%s/%s" % (save_dir, split_dir, filenames[i])
            mkdir_p(os.path.dirname(stem))
            tiles = torch.stack([imgs[i].view(3, imsize, imsize) for imgs in images_list])
            Image.fromarray(image_grid_uint8(tiles, nrow=10).cpu().numpy()).save("%s_%d.png" % (stem, imsize))

    def _eval_save_dir(self):
        """cub:829-844 naming: eval/Testset_<mode>_fixednoise[_clsprior-random]_<epoch tag>_<run directory>."""
        path = cfg.TRAIN.NET_G
        tag = os.path.splitext(path)[0].split("_")[-1]
        parts = path.split("/")
        run = parts[-3] if len(parts) >= 3 else "run"
        mode = "evalmode" if cfg.TEST.EVAL_MODE else "trainmode"
        prior = "_clsprior-random" if cfg.TEST.CLS_PRIOR else ""
        return "eval/Testset_%s_fixednoise%s_%s_%s" % (mode, prior, tag, run)

    MAX_EVAL_SENTENCES = 10            # cub:826 embedding_dim

    def evaluate(self, split_dir, save_dir=None):
        """cub:776-911: load cfg.TRAIN.NET_G into a freshly built generator, and for every batch of the test loader
        `(imgs, t_embeddings[B, n_sentences, T], cls (1-based), keys)` draw ONE noise batch, generate the final-stage
        image for each of the first 10 sentence embeddings, and write tile sheets (cfg.TEST.B_EXAMPLE) or single PNGs.
        BatchNorm runs on the running statistics when cfg.TEST.EVAL_MODE (the kernels' eval-mode normalisation).
        Returns the number of images written."""
        if cfg.TRAIN.NET_G == "":
            print("Error: cfg.TRAIN.NET_G is empty, no generator snapshot to evaluate")
            return 0
        netG, _ = build_G(use_cap=cfg.TEST.G_CAPSULE)
        netG.load_state_dict(_strip_module(torch.load(cfg.TRAIN.NET_G, map_location="cpu")))
        netG.to(self.device)
        model.to_kernel_layout(netG)
        netG.eval() if cfg.TEST.EVAL_MODE else netG.train()
        save_dir = save_dir or self._eval_save_dir()
        final = cfg.TREE.BASE_SIZE * cfg.TREE.SCALE ** (cfg.TREE.BRANCH_NUM - 1)
        split_z = isinstance(netG, model.COND_G_NET_CATZ_CA)
        from .engine import tensor_core_matmul
        count = 0
        for step, data in enumerate(self.data_loader, 0):
            imgs, t_embeddings, cls, filenames = data
            cls = cls.long() - 1
            emb = t_embeddings.to(self.device, non_blocking=True)
            cls_onehot = self.onehot(cls.to(self.device), cfg.GAN.ENTITY_DIM)
            B = emb.shape[0]
            noise = torch.randn(B, cfg.GAN.Z_DIM, device=self.device)           # one draw per batch (cub:866-867)
            sheets = []
            with torch.no_grad(), tensor_core_matmul():
                for i in range(min(self.MAX_EVAL_SENTENCES, emb.shape[1])):
                    sen = emb[:, i, :].contiguous()
                    if not split_z:
                        hcodes = netG(noise, sen)[0]                            # USE_CLS False branch (cub:135)
                    elif cfg.TEST.CLS_PRIOR:
                        hcodes = netG(noise, sen)[0]                            # class code from the N(0,1) prior (model.py:491-495)
                    else:
                        hcodes = netG(noise, sen, cls_onehot)[0]
                    fake = netG.image(hcodes)[-1]
                    if cfg.TEST.B_EXAMPLE:
                        sheets.append(fake)
                    else:
                        self.save_singleimages(fake, filenames, save_dir, split_dir, i, cls, final, 0)
                    count += B
            if cfg.TEST.B_EXAMPLE:
                self.save_superimages(sheets, filenames, save_dir, split_dir, final)
            print("[%d/%d]" % (step, self.num_batches))
        print("Number of images: %d -> %s" % (count, save_dir))
        return count
