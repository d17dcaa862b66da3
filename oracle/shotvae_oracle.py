"""CPU oracle for the SHOT-VAE / M2-VAE training step.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package (`shot-vae_b200/`) never does.

What it is: a *functional* restatement, in plain torch FP32 on the CPU, of the arithmetic the
reference performs on its hot path.  The reference's arithmetic lives in PyTorch (third party, pinned
only in prose: torch 1.2.0 in reference README.md:17-23; torch 2.11.0 in this image), so the oracle
states the same ATen ops over a flat `state` dict that uses the reference's own state_dict key names.
Each function cites the reference file:line it follows.

Parity status: the reference has no tests/golden vectors of its own.  The oracle is
pinned against outputs of the reference itself executed in the build container
(tests/golden/make_golden.py -> tests/golden/*.json, checked by tests/test_oracle_golden.py).
"""
import math
import re
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# network topology (reference shot_vae_model/wideresnet.py:68-114, preactresnet.py:85-133)
# ----------------------------------------------------------------------------------------------
def encoder_topology(name):
    """Returns dict(init_features, units=[(prefix, cin, cout, stride, has_shortcut)], slope,
    shortcut_act, feat)."""
    if "wideresnet" in name:
        depth, width = [int(v) for v in re.findall(r"\d+", name)]          # wideresnet.py:110-112
        assert (depth - 4) % 6 == 0, "depth should be 6n+4"                 # wideresnet.py:72
        block_depth = (depth - 4) // 6
        widths = [int(v * width) for v in (16, 32, 64)]                    # wideresnet.py:74
        units, cin = [], 16
        for b, w in enumerate(widths):
            for u in range(block_depth):
                stride = 2 if (u == 0 and b > 0) else 1                     # wideresnet.py:57-58,83
                ci = cin if u == 0 else w
                units.append(("feature_extractor.encoder.wideblock%d.wide_block.wideunit%d" % (b + 1, u + 1),
                              ci, w, stride, ci != w or stride != 1))       # wideresnet.py:37
            cin = w
        return dict(init_features=16, units=units, slope=0.01, shortcut_act=True, feat=widths[-1])
    if name == "preactresnet18":
        cfg = [2, 2, 2, 2]                                                  # preactresnet.py:121
        units, cin, cout = [], 64, 64
        for b, depth in enumerate(cfg):
            for u in range(depth):
                stride = 2 if (u == 0 and b > 0) else 1                     # preactresnet.py:73-74,101
                ci = cin if u == 0 else cout
                units.append(("feature_extractor.encoder.block%d.preact_block.unit%d" % (b + 1, u + 1),
                              ci, cout, stride, stride != 1 or ci != cout))  # preactresnet.py:52
            cin, cout = cout, cout * 2
        return dict(init_features=64, units=units, slope=0.0, shortcut_act=False, feat=512)
    raise NotImplementedError("{} not implemented".format(name))           # vae.py:106


DEC_PLAN = [(1024, None), (512, 4), (256, 4), (128, 4), (64, 4)]            # decoder.py:12-57 (num_feature=64)


def _conv_init(shape, bias):
    """torch.nn.modules.conv._ConvNd.reset_parameters: kaiming_uniform(a=sqrt(5)) + uniform bias."""
    w = torch.empty(shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = None
    if bias:
        fan_in = shape[1] * int(np.prod(shape[2:]))
        bound = 1 / math.sqrt(fan_in)
        b = torch.empty(shape[0]).uniform_(-bound, bound)
    return w, b


def _bn_init(state, prefix, c):
    state[prefix + ".weight"] = torch.ones(c)
    state[prefix + ".bias"] = torch.zeros(c)
    state[prefix + ".running_mean"] = torch.zeros(c)
    state[prefix + ".running_var"] = torch.ones(c)
    state[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def init_state(encoder_name, nd, ldc=128, in_ch=3, seed=1):
    """Default-constructor initialisation in the reference's module construction order
    (vae.py:92-138, wideresnet.py:26-43, decoder.py:11-62) so that, under the same torch seed, the
    tensors are bit-identical to `VariationalAutoEncoder(..., data_parallel=False)`."""
    topo = encoder_topology(encoder_name)
    torch.manual_seed(seed)
    st = OrderedDict()
    f0 = topo["init_features"]
    w, b = _conv_init((f0, in_ch, 3, 3), True)
    st["feature_extractor.encoder.pre_process.conv0.weight"] = w
    st["feature_extractor.encoder.pre_process.conv0.bias"] = b
    for prefix, ci, co, stride, sc in topo["units"]:
        _bn_init(st, prefix + ".f_block.norm1", ci)
        st[prefix + ".f_block.conv1.weight"] = _conv_init((co, ci, 3, 3), False)[0]
        _bn_init(st, prefix + ".f_block.norm2", co)
        st[prefix + ".f_block.conv2.weight"] = _conv_init((co, co, 3, 3), False)[0]
        if sc:
            _bn_init(st, prefix + ".i_block.norm", ci)
            st[prefix + ".i_block.conv.weight"] = _conv_init((co, ci, 1, 1), False)[0]
    _bn_init(st, "feature_extractor.encoder.transition.norm", topo["feat"])
    for head, n in (("continuous_inference.mean", ldc), ("continuous_inference.log_sigma", ldc),
                    ("disc_latent_inference", nd)):
        w = torch.empty(n, topo["feat"])
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(topo["feat"])
        st[head + ".fc.weight"] = w
        st[head + ".fc.bias"] = torch.empty(n).uniform_(-bound, bound)
    cin, idx = ldc + nd, 0
    for cout, k in DEC_PLAN:
        kk = 1 if k is None else k
        st["feature_reconstructor.decoder.%d.weight" % idx] = _conv_init((cin, cout, kk, kk), False)[0]
        _bn_init(st, "feature_reconstructor.decoder.%d" % (idx + 1), cout)
        cin, idx = cout, idx + 3
    st["feature_reconstructor.decoder.%d.weight" % idx] = _conv_init((cin, in_ch, 4, 4), False)[0]
    return st


def param_names(state):
    return [k for k in state if not (k.endswith("running_mean") or k.endswith("running_var")
                                     or k.endswith("num_batches_tracked"))]


# ----------------------------------------------------------------------------------------------
# host RNG draws (reference order: SURVEY.md section 7 item 4)
# ----------------------------------------------------------------------------------------------
class LiveDraws:
    """Draws from the torch CPU generator / numpy global RNG exactly where the reference does
    (vae.py:69,82; mixup.py:7,21,31,35) and records them."""

    def __init__(self):
        self.log = []

    def randn(self, *shape):
        t = torch.randn(*shape); self.log.append(("randn", t)); return t

    def rand(self, *shape):
        t = torch.rand(*shape); self.log.append(("rand", t)); return t

    def beta(self, a, b):
        v = float(np.random.beta(a, b)); self.log.append(("beta", v)); return v

    def randperm(self, n):
        t = torch.randperm(n); self.log.append(("randperm", t)); return t


class ReplayDraws:
    def __init__(self, log):
        self.log, self.pos = list(log), 0

    def _next(self, kind):
        k, v = self.log[self.pos]; self.pos += 1
        assert k == kind, "draw order mismatch: wanted %s got %s" % (kind, k)
        return v

    def randn(self, *shape): return self._next("randn")
    def rand(self, *shape): return self._next("rand")
    def beta(self, a, b): return self._next("beta")
    def randperm(self, n): return self._next("randperm")


# ----------------------------------------------------------------------------------------------
# forward pieces
# ----------------------------------------------------------------------------------------------
def _bn(st, prefix, x, training):
    if training:
        st[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, st[prefix + ".running_mean"], st[prefix + ".running_var"],
                        st[prefix + ".weight"], st[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def _act(x, slope):
    return F.leaky_relu(x, slope) if slope != 0.0 else F.relu(x)


def encoder_forward(st, topo, x, training=True):
    """wideresnet.py:45-49,97-99 / preactresnet.py:61-65,114-116."""
    slope = topo["slope"]
    h = F.conv2d(x, st["feature_extractor.encoder.pre_process.conv0.weight"],
                 st["feature_extractor.encoder.pre_process.conv0.bias"], 1, 1)
    for prefix, ci, co, stride, sc in topo["units"]:
        f = _act(_bn(st, prefix + ".f_block.norm1", h, training), slope)
        f = F.conv2d(f, st[prefix + ".f_block.conv1.weight"], None, stride, 1)
        f = _act(_bn(st, prefix + ".f_block.norm2", f, training), slope)     # Dropout(p=0) is identity
        f = F.conv2d(f, st[prefix + ".f_block.conv2.weight"], None, 1, 1)
        if sc:
            s = _bn(st, prefix + ".i_block.norm", h, training)
            if topo["shortcut_act"]:
                s = _act(s, slope)
            h = F.conv2d(s, st[prefix + ".i_block.conv.weight"], None, stride, 0)
        h = f + h
    return _act(_bn(st, "feature_extractor.encoder.transition.norm", h, training), slope)


def heads_forward(st, feat):
    """vae.py:10-15,143-146."""
    mu = F.linear(feat, st["continuous_inference.mean.fc.weight"], st["continuous_inference.mean.fc.bias"])
    ls = F.linear(feat, st["continuous_inference.log_sigma.fc.weight"], st["continuous_inference.log_sigma.fc.bias"])
    la = F.log_softmax(F.linear(feat, st["disc_latent_inference.fc.weight"], st["disc_latent_inference.fc.bias"]), dim=1)
    return mu, ls, la


def sample_latent(mu, ls, la, draws, temperature, disc_label=None, mixup=False, disc_label_mixup=None,
                  mixup_lam=None):
    """vae.py:23-86."""
    z = mu + torch.exp(ls) * draws.randn(*mu.shape)                          # vae.py:76-86
    if disc_label is not None:
        c = torch.zeros(la.shape).scatter(1, disc_label.view(-1, 1), 1)      # vae.py:42-49
        if mixup:
            cb = torch.zeros(la.shape).scatter(1, disc_label_mixup.view(-1, 1), 1)
            c = mixup_lam * c + (1 - mixup_lam) * cb
    else:
        eps = 1e-12                                                          # vae.py:68-73
        unif = draws.rand(*la.shape)
        gumbel = -torch.log(-torch.log(unif + eps) + eps)
        c = torch.softmax((la + gumbel) / temperature, dim=1)
    lat = torch.cat([z, c], dim=1)
    return lat.view(lat.size(0), lat.size(1), 1, 1)


def decoder_forward(st, lat, training=True):
    """decoder.py:11-69 (kernel_size = img/32 = 1 for 32x32 inputs, vae.py:134)."""
    h, idx = lat, 0
    for cout, k in DEC_PLAN:
        w = st["feature_reconstructor.decoder.%d.weight" % idx]
        h = F.conv_transpose2d(h, w, None, 1, 0) if k is None else F.conv_transpose2d(h, w, None, 2, 1)
        h = F.relu(_bn(st, "feature_reconstructor.decoder.%d" % (idx + 1), h, training))
        idx += 3
    return F.conv_transpose2d(h, st["feature_reconstructor.decoder.%d.weight" % idx], None, 2, 1)


def vae_forward(st, topo, x, draws, temperature=0.67, mixup=False, disc_label=None, disc_pseudo_label=None,
                mixup_lam=None, training=True):
    """vae.py:140-151 -> (reconstruction, norm_mean, norm_log_sigma, disc_log_alpha)."""
    feat = encoder_forward(st, topo, x, training)
    feat = F.adaptive_avg_pool2d(feat, (1, 1)).view(x.size(0), -1)
    mu, ls, la = heads_forward(st, feat)
    lat = sample_latent(mu, ls, la, draws, temperature, disc_label, mixup, disc_pseudo_label, mixup_lam)
    return decoder_forward(st, lat, training), mu, ls, la


# ----------------------------------------------------------------------------------------------
# criteria (lib/criterion.py) and mixup (lib/utils/mixup.py)
# ----------------------------------------------------------------------------------------------
def vae_criterion(x, x_rec, mu, ls, la, nd, x_sigma=1.0, bce=True):
    """criterion.py:32-57."""
    b = x.size(0)
    if bce:
        rec = F.binary_cross_entropy_with_logits(x_rec, x, reduction="sum") / b
    else:
        rec = F.mse_loss(torch.sigmoid(x_rec), x, reduction="sum") / (2 * b * (x_sigma ** 2))
    ls2 = 2 * ls
    klc = 0.5 * torch.sum(mu * mu + torch.exp(ls2) - ls2 - 1) / b
    log_prior = torch.log(torch.tensor([1 / nd for _ in range(nd)]).view(1, -1).float())   # criterion.py:29-30
    kld = torch.sum(torch.exp(la) * (la - log_prior)) / b
    return rec, klc, kld


def cls_criterion(predict, label, batch_weight=None):
    """criterion.py:97-108."""
    if batch_weight is None:
        return -1 * torch.mean(torch.sum(predict * label, dim=1))
    return -1 * torch.mean(torch.sum(predict * label, dim=1) * batch_weight)


def gaussian_kl(mu1, ls1, mu2, ls2):
    """mixup.py:93-99 (one pair)."""
    dim = mu1.size(0)
    s1, s2 = torch.exp(ls1), torch.exp(ls2)
    return torch.sum(ls2 - ls1) + 0.5 * torch.sum(s1 ** 2 / s2 ** 2) + 0.5 * torch.sum(
        (mu1 - mu2) ** 2 / (s2 ** 2)) - 0.5 * dim


def pairwise_kl_matrix(mu, ls):
    """mixup.py:11-16: kl[i, j] = KL(N_i || N_j), same per-pair operation order (vectorised over j)."""
    b = mu.size(0)
    kl = torch.zeros(b, b)
    for i in range(b):
        s1 = torch.exp(ls[i])
        s2 = torch.exp(ls)
        kl[i] = torch.sum(ls - ls[i], dim=1) + 0.5 * torch.sum(s1 ** 2 / s2 ** 2, dim=1) + 0.5 * torch.sum(
            (mu[i] - mu) ** 2 / (s2 ** 2), dim=1) - 0.5 * mu.size(1)
    return kl


def optimal_match_index(mu, ls):
    """mixup.py:17-18: second entry of the ascending top-2 of each row."""
    _, index = torch.topk(pairwise_kl_matrix(mu, ls), 2, largest=False)
    return index[:, 1]


def mixup_vae_data(image, mu, ls, la, draws, optimal_match=False):
    """mixup.py:5-26."""
    lam = draws.beta(2.0, 2.0)
    index = optimal_match_index(mu, ls) if optimal_match else draws.randperm(image.size(0))
    mixed_image = lam * image + (1 - lam) * image[index, :]
    mixed_mu = lam * mu + (1 - lam) * mu[index]
    mixed_sigma = lam * torch.exp(ls) + (1 - lam) * torch.exp(ls[index])
    mixed_alpha = lam * torch.exp(la) + (1 - lam) * torch.exp(la[index])
    return mixed_image, mixed_mu, mixed_sigma, mixed_alpha, lam, index


def label_smoothing(image, mu, ls, la, draws, epsilon=0.1, disc_label=None):
    """mixup.py:29-41."""
    lam = draws.beta(epsilon, epsilon) if epsilon > 0 else 1
    index = draws.randperm(image.size(0))
    s_image = lam * image + (1 - lam) * image[index, :]
    s_mu = lam * mu + (1 - lam) * mu[index]
    s_sigma = lam * torch.exp(ls) + (1 - lam) * torch.exp(ls[index])
    s_alpha = lam * torch.exp(la) + (1 - lam) * torch.exp(la[index])
    return s_image, s_mu, s_sigma, s_alpha, disc_label[index], lam, index


# ----------------------------------------------------------------------------------------------
# the training step (main_shot_vae.py:261-366, main_M2_vae.py:242-307)
# ----------------------------------------------------------------------------------------------
def alpha_schedule(epoch, max_epoch, alpha_max):
    """main_shot_vae.py:518-520."""
    return alpha_max * math.exp(-5 * (1 - min(1, epoch / max_epoch)) ** 2)


def default_hyper(dataset="Cifar10", m2=False):
    """argparse defaults main_shot_vae.py:30-106 + dataset overrides :139,:161-163
    (main_M2_vae.py:123,146-147 for M2)."""
    h = dict(epochs=600, akb=200, aew=400, apw=200, ewm=1e-3, kbmc=1e-3, kbmd=1e-3, pwm=1.0, wrd=1.0,
             wmf=0.4, cmi=0.0, dmi=2.3, epsilon=0.1, om=False, lr=0.1, momentum=0.9, wd=5e-4,
             temperature=0.67, x_sigma=1.0, br=True)
    if dataset == "Cifar100":
        h.update(akb=150, apw=400, dmi=4.6)
    if m2:
        h.update(cmi=200.0 if dataset == "Cifar10" else 1280.0)
    return h


def schedules(h, epoch):
    """main_shot_vae.py:270-279."""
    return dict(cmi=alpha_schedule(epoch, h["akb"], h["cmi"]), dmi=alpha_schedule(epoch, h["akb"], h["dmi"]),
                ew=alpha_schedule(epoch, h["aew"], h["ewm"]), kbc=alpha_schedule(epoch, h["akb"], h["kbmc"]),
                kbd=alpha_schedule(epoch, h["akb"], h["kbmd"]), pwm=alpha_schedule(epoch, h["apw"], h["pwm"]),
                ucw=alpha_schedule(epoch, round(h["wmf"] * h["epochs"]), h["wrd"]))


def _f(t):
    return float(t.detach()) if torch.is_tensor(t) else float(t)


def _onehot(y, n):
    return torch.zeros(y.size(0), n).scatter_(1, y.view(-1, 1), 1)


def _require_grad(st):
    for k in param_names(st):
        st[k].requires_grad_(True)


def shot_step(st, encoder_name, nd, image_l, label_l, image_u, label_u, epoch, hyper, draws, keep=False):
    """One (labelled, unlabelled) iteration of main_shot_vae.train, :281-364, WITHOUT the optimizer
    step.  Gradients are accumulated in `st[k].grad`.  Returns a dict of every loss term (Python
    floats) and, with keep=True, the intermediate tensors."""
    topo = encoder_topology(encoder_name)
    s = schedules(hyper, epoch)
    T, bce, xs = hyper["temperature"], hyper["br"], hyper["x_sigma"]
    _require_grad(st)
    out = {}
    bl, bu = image_l.size(0), image_u.size(0)
    onehot_l = _onehot(label_l, nd)
    # P1 (:288-296)
    rec_l, mu_l, ls_l, la_l = vae_forward(st, topo, image_l, draws, T, disc_label=label_l)
    rl, kc, kd = vae_criterion(image_l, rec_l, mu_l, ls_l, la_l, nd, xs, bce)
    prior_l = s["kbc"] * torch.abs(kc - s["cmi"]) + s["kbd"] * torch.abs(kd - s["dmi"])
    elbo_l = rl + prior_l
    with torch.no_grad():                                                    # :297-310
        s_img, s_mu, s_sig, s_alpha, s_lab, lam_l, idx_l = label_smoothing(
            image_l, mu_l, ls_l, la_l, draws, hyper["epsilon"], label_l)
        s_onehot = _onehot(s_lab, nd)
    # P2 (:311-324)
    rec2, mu2, ls2, la2 = vae_forward(st, topo, s_img, draws, T, True, label_l, s_lab, lam_l)
    disc_post_l = lam_l * cls_criterion(la2, onehot_l) + (1 - lam_l) * cls_criterion(la2, s_onehot)
    cont_post_l = (F.mse_loss(mu2, s_mu, reduction="sum") + F.mse_loss(torch.exp(ls2), s_sig, reduction="sum")) / bl
    elbo_l = elbo_l + s["kbc"] * s["pwm"] * cont_post_l
    loss_sup = s["ew"] * elbo_l + disc_post_l
    loss_sup.backward()
    # P3 (:326-346)
    rec_u, mu_u, ls_u, la_u = vae_forward(st, topo, image_u, draws, T)
    with torch.no_grad():                                                    # :331-339
        lsu = torch.zeros(bu, nd).scatter_(1, label_u.view(-1, 1), 1 - 0.001 - 0.001 / (nd - 1))
        lsu = lsu + torch.ones(lsu.size()) * 0.001 / (nd - 1)
        au = torch.exp(la_u)
        kl_inf = float(torch.sum(au * la_u - au * torch.log(lsu)) / bu)
    ru, kcu, kdu = vae_criterion(image_u, rec_u, mu_u, ls_u, la_u, nd, xs, bce)
    prior_u = s["kbc"] * torch.abs(kcu - s["cmi"]) + s["kbd"] * torch.abs(kdu - s["dmi"])
    elbo_u = ru + prior_u
    with torch.no_grad():                                                    # :348-355
        m_img, m_mu, m_sig, m_alpha, lam_u, idx_u = mixup_vae_data(image_u, mu_u, ls_u, la_u, draws, hyper["om"])
    # P4 (:356-364)
    rec4, mu4, ls4, la4 = vae_forward(st, topo, m_img, draws, T)
    disc_post_u = cls_criterion(la4, m_alpha)
    cont_post_u = (F.mse_loss(mu4, m_mu, reduction="sum") + F.mse_loss(torch.exp(ls4), m_sig, reduction="sum")) / bu
    elbo_u = elbo_u + s["kbc"] * s["pwm"] * cont_post_u
    loss_unsup = s["ew"] * elbo_u + s["ucw"] * disc_post_u
    loss_unsup.backward()
    out.update(rec_l=_f(rl), klc_l=_f(kc), kld_l=_f(kd), prior_l=_f(prior_l), cont_post_l=_f(cont_post_l),
               disc_post_l=_f(disc_post_l), loss_sup=_f(loss_sup), lam_l=_f(lam_l),
               rec_u=_f(ru), klc_u=_f(kcu), kld_u=_f(kdu), prior_u=_f(prior_u), cont_post_u=_f(cont_post_u),
               disc_post_u=_f(disc_post_u), loss_unsup=_f(loss_unsup), lam_u=_f(lam_u), kl_inference=kl_inf)
    if keep:
        out["tensors"] = dict(rec_l=rec_l, mu_l=mu_l, ls_l=ls_l, la_l=la_l, idx_l=idx_l, s_img=s_img, s_mu=s_mu,
                              s_sig=s_sig, s_alpha=s_alpha, s_lab=s_lab, rec2=rec2, mu2=mu2, ls2=ls2, la2=la2,
                              rec_u=rec_u, mu_u=mu_u, ls_u=ls_u, la_u=la_u, idx_u=idx_u, m_img=m_img, m_mu=m_mu,
                              m_sig=m_sig, m_alpha=m_alpha, rec4=rec4, mu4=mu4, ls4=ls4, la4=la4)
    return out


def m2_step(st, encoder_name, nd, image_l, label_l, image_u, label_u, epoch, hyper, draws, keep=False):
    """main_M2_vae.train :258-305 without the optimizer step."""
    topo = encoder_topology(encoder_name)
    s = schedules(hyper, epoch)
    T, bce, xs = hyper["temperature"], hyper["br"], hyper["x_sigma"]
    _require_grad(st)
    b = min(image_l.size(0), image_u.size(0))                                # :259-266
    image_l, label_l, image_u, label_u = image_l[:b], label_l[:b], image_u[:b], label_u[:b]
    onehot_l = _onehot(label_l, nd)
    rec_l, mu_l, ls_l, la_l = vae_forward(st, topo, image_l, draws, T, disc_label=label_l)
    rl, kc, kd = vae_criterion(image_l, rec_l, mu_l, ls_l, la_l, nd, xs, bce)
    prior_l = s["kbc"] * torch.abs(kc - s["cmi"]) + s["kbd"] * torch.abs(kd - s["dmi"])
    disc_post_l = cls_criterion(la_l, onehot_l)
    loss_sup = s["ew"] * (rl + prior_l) + disc_post_l
    loss_sup.backward()
    rec_u, mu_u, ls_u, la_u = vae_forward(st, topo, image_u, draws, T)
    with torch.no_grad():
        lsu = torch.zeros(b, nd).scatter_(1, label_u.view(-1, 1), 1 - 0.001 - 0.001 / (nd - 1))
        lsu = lsu + torch.ones(lsu.size()) * 0.001 / (nd - 1)
        au = torch.exp(la_u)
        kl_inf = float(torch.sum(au * la_u - au * torch.log(lsu)) / b)
    ru, kcu, kdu = vae_criterion(image_u, rec_u, mu_u, ls_u, la_u, nd, xs, bce)
    prior_u = s["kbc"] * torch.abs(kcu - s["cmi"]) + s["kbd"] * torch.abs(kdu - s["dmi"])
    loss_unsup = s["ew"] * (ru + prior_u)
    loss_unsup.backward()
    out = dict(rec_l=_f(rl), klc_l=_f(kc), kld_l=_f(kd), prior_l=_f(prior_l), disc_post_l=_f(disc_post_l),
               loss_sup=_f(loss_sup), rec_u=_f(ru), klc_u=_f(kcu), kld_u=_f(kdu), prior_u=_f(prior_u),
               loss_unsup=_f(loss_unsup), kl_inference=kl_inf)
    if keep:
        out["tensors"] = dict(rec_l=rec_l, mu_l=mu_l, ls_l=ls_l, la_l=la_l, rec_u=rec_u, mu_u=mu_u, ls_u=ls_u, la_u=la_u)
    return out


def sgd_step(st, momentum_buf, lr, momentum=0.9, wd=5e-4):
    """torch.optim.SGD(momentum, weight_decay) on every parameter (main_shot_vae.py:198,365-366);
    first step initialises the buffer with the gradient (torch semantics)."""
    with torch.no_grad():
        for k in param_names(st):
            p = st[k]
            if p.grad is None:
                continue
            g = p.grad + wd * p
            if k not in momentum_buf:
                momentum_buf[k] = g.clone()
            else:
                momentum_buf[k].mul_(momentum).add_(g)
            p.add_(momentum_buf[k], alpha=-lr)
            p.grad = None


def clone_state(st):
    return OrderedDict((k, v.detach().clone()) for k, v in st.items())


def synthetic_batch(batch, nd, seed):
    """SURVEY.md section 8d synthetic inputs: images U[0,1), labels randint."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 3, 32, 32, generator=g), torch.randint(0, nd, (batch,), generator=g),
            torch.rand(batch, 3, 32, 32, generator=g), torch.randint(0, nd, (batch,), generator=g))


# ----------------------------------------------------------------------------------------------
# entry points the reference defines but never calls
# ----------------------------------------------------------------------------------------------
def _rec_term(x, x_rec, x_sigma, bce):
    B = x.size(0)
    if bce:                                                               # criterion.py:67-68 / 125-126
        return F.binary_cross_entropy_with_logits(x_rec, x, reduction="sum") / B
    return F.mse_loss(torch.sigmoid(x_rec), x, reduction="sum") / (2 * B * x_sigma ** 2)   # criterion.py:70-71 / 128-129


def _kl_std_normal(mu, ls):
    return 0.5 * torch.sum(mu * mu + torch.exp(2 * ls) - 2 * ls - 1) / mu.size(0)           # criterion.py:73-76


def m1_criterion(x, x_rec, mu, ls, x_sigma=1.0, bce=True):
    """M1Criterion.forward (criterion.py:64-76)"""
    return _rec_term(x, x_rec, x_sigma, bce), _kl_std_normal(mu, ls)


def m2_criterion(mu, ls, la, nd):
    """M2Criterion.forward (criterion.py:83-91)"""
    prior = torch.log(torch.tensor([1 / nd for _ in range(nd)]).view(1, -1).float())
    return _kl_std_normal(mu, ls), torch.sum(torch.exp(la) * (la - prior)) / mu.size(0)


def reconstruction_criterion(x, x_rec, x_sigma=1.0, bce=True):
    """ReconstructionCriterion.forward (criterion.py:122-131)"""
    return _rec_term(x, x_rec, x_sigma, bce)


def kl_norm_criterion(mu_pre, ls_pre, mu_gt=None, sigma_gt=None):
    """KLNormCriterion.forward (criterion.py:138-158)"""
    if mu_gt is None or sigma_gt is None:
        return _kl_std_normal(mu_pre, ls_pre)
    v_gt = sigma_gt ** 2
    return 0.5 * torch.sum(2 * torch.log(sigma_gt + 1e-4) - 2 * ls_pre + torch.exp(2 * ls_pre) / v_gt
                           + (mu_pre - mu_gt) ** 2 / v_gt - 1) / mu_pre.size(0)


def kl_disc_criterion(log_pre, gt, qp_order=True):
    """KLDiscCriterion.forward (criterion.py:170-177)"""
    log_gt = torch.log(gt + 1e-4)
    if qp_order:
        return torch.sum(torch.exp(log_pre) * (log_pre - log_gt)) / log_pre.size(0)
    return torch.sum(gt * (log_gt - log_pre)) / log_pre.size(0)


def pairwise_norm_kl_dist(u1, ls1, u2, ls2):
    """pairwise_norm_kl_dist_gpu (calculate_dist.py:94-107)"""
    v1, v2 = torch.exp(ls1) ** 2, torch.exp(ls2) ** 2
    ratio = v1.unsqueeze(1) / v2.unsqueeze(0)
    shift = (u1.unsqueeze(1) - u2.unsqueeze(0)) ** 2 / v2.unsqueeze(0)
    return 0.5 * (-torch.log(ratio).sum(2) + ratio.sum(2) + shift.sum(2) - ls1.size(1))


def pairwise_square_euclidean(v1, v2):
    """pairwise_square_euclidean_gpu (calculate_dist.py:110-117)"""
    return ((v1.unsqueeze(1) - v2.unsqueeze(0)) ** 2).sum(2)


def pairwise_norm_wasserstein_dist(u1, ls1, u2, ls2):
    """pairwise_norm_wasserstein_dist_gpu (calculate_dist.py:120-130)"""
    return pairwise_square_euclidean(u1, u2) + pairwise_square_euclidean(torch.exp(ls1), torch.exp(ls2))


def mean_dist_pairwise(u1, u2, distance="euclidean"):
    """calculate_mean_dist_pairwise (calculate_dist.py:133-160); "cosine" divides by squared norms as written"""
    if distance == "euclidean":
        return pairwise_square_euclidean(u1, u2)
    if distance == "cosine":
        return torch.mm(u1, u2.t()) / ((u1 ** 2).sum(1).view(-1, 1) * (u2 ** 2).sum(1).view(1, -1))
    raise NotImplementedError("distance {} not implemented".format(distance))


# ----------------------------------------------------------------------------------------------
# input pipeline (reference lib/dataloader.py)
# ----------------------------------------------------------------------------------------------
def augment_batch(data, index, params, pad=4, out_size=32, hwc=True):
    """The reference's train transform Pad(pad, reflect) -> RandomHorizontalFlip -> RandomCrop(out_size) -> ToTensor
    (dataloader.py:42-70) for given per-sample parameters params[b] = (crop row, crop column, flip); params None =
    the test transform (ToTensor only).  data: uint8 numpy [N, H, W, C] (hwc) or [N, C, H, W]; -> float32 [B, C, out, out]."""
    data = np.asarray(data)
    outs = []
    for b, i in enumerate(np.asarray(index).tolist()):
        img = data[i] if hwc else np.transpose(data[i], (1, 2, 0))        # H, W, C
        ci, cj, flip = (0, 0, 0) if params is None else [int(v) for v in params[b]]
        p = pad if params is not None else 0
        img = np.pad(img, ((p, p), (p, p), (0, 0)), mode="reflect") if p else img   # transforms.Pad(padding_mode='reflect')
        if flip:
            img = img[:, ::-1]                                               # RandomHorizontalFlip on the padded image
        img = img[ci:ci + out_size, cj:cj + out_size]                        # RandomCrop
        outs.append(np.transpose(img, (2, 0, 1)).astype(np.float32) / np.float32(255.0))   # ToTensor
    return torch.from_numpy(np.stack(outs))


def per_class_split(labels, num_classes, valid_num_per_class, annotated_num_per_class=None):
    """get_ssl_sampler / get_cifar10_sl_sampler index lists (dataloader.py:73-92,115-139): per class one
    torch.randperm over its positions; valid = first v, labelled = next a, unlabelled / train = everything after v.
    -> (valid, train_l, train_u) or (valid, train) when annotated_num_per_class is None."""
    v, a = valid_num_per_class, annotated_num_per_class
    valid, lab, rest = [], [], []
    for c in range(num_classes):
        loc = torch.nonzero(labels == c).view(-1)
        loc = loc[torch.randperm(loc.size(0))]
        valid.extend(loc[:v].tolist())
        if a is not None:
            lab.extend(loc[v:v + a].tolist())
        rest.extend(loc[v:].tolist())
    return (valid, rest) if a is None else (valid, lab, rest)
