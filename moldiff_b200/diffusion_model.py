"""``MolDiff`` -- the joint (atom type, position, bond type) diffusion model, B200-native.

Drop-in for ``models.model.MolDiff`` (reference ``models/model.py:12-378``): same constructor
``MolDiff(config, num_node_types, num_edge_types)``, same ``forward`` / ``sample`` / ``get_loss`` /
``add_noise`` signatures and return structures, same ``state_dict`` schema (581 keys).  What differs is
where the arithmetic runs: ``forward`` is ONE C-ABI call (``mdb_moldiff_forward``: embedders + 6
NodeEdgeNet blocks + decoders as hand-written sm_100a kernels); the guidance term of ``sample`` is the
bond predictor's CUDA forward + hand-written input-gradient backward.  The noise schedule / posterior
sampling around it stays host-side PyTorch (north_star), operating on device tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine, train_path
from .nets import MLP, GaussianSmearing, NodeEdgeNet
from .schedules import get_beta_schedule
from .transitions import CategoricalTransition, GaussianTransition, gumbel_argmax


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def build_transitions(owner, diff_cfg, num_node_types, num_edge_types, with_edges=True):
    """Shared by MolDiff and BondPredictor (model.py:49-95, bond_predictor.py:40-73)."""
    owner.num_timesteps = diff_cfg["num_timesteps"]
    owner.categorical_space = _cfg_get(diff_cfg, "categorical_space", "discrete")
    if owner.categorical_space == "continuous":
        owner.scaling = _cfg_get(diff_cfg, "scaling", [1.0, 1.0, 1.0])
    else:
        owner.scaling = [1.0, 1.0, 1.0]
    if owner.scaling[0] != 1:
        raise AssertionError("scaling for pos should be 1")
    T = owner.num_timesteps
    owner.pos_transition = GaussianTransition(get_beta_schedule(num_timesteps=T, **diff_cfg["diff_pos"]))
    specs = [("node_transition", "diff_atom", num_node_types, 1)]
    if with_edges:
        specs.append(("edge_transition", "diff_bond", num_edge_types, 2))
    for attr, key, k, si in specs:
        betas = get_beta_schedule(num_timesteps=T, **diff_cfg[key])
        if owner.categorical_space == "discrete":
            tr = CategoricalTransition(betas, k, init_prob=diff_cfg[key]["init_prob"])
        elif owner.categorical_space == "continuous":
            tr = GaussianTransition(betas, k, owner.scaling[si])
        else:
            raise ValueError(owner.categorical_space)
        setattr(owner, attr, tr)


class _PackedMixin:
    """Lazily packs the module's weights into the kernel layout; re-packs when any parameter was
    modified in place (optimizer step, load_state_dict) or moved (.to())."""

    def _packed_net(self, device):
        params = list(self.parameters()) + list(self.buffers())
        key = (str(device), tuple(p._version for p in params), tuple(p.data_ptr() for p in params))
        if getattr(self, "_packed", None) is None or self._packed_key != key:
            self._packed = self._pack(device)
            self._packed_key = key
        return self._packed


class MolDiff(nn.Module, _PackedMixin):
    def __init__(self, config, num_node_types, num_edge_types, **kwargs):
        super().__init__()
        self.config = config
        self.num_node_types = num_node_types
        self.num_edge_types = num_edge_types
        self.bond_len_loss = _cfg_get(config, "bond_len_loss", False)
        build_transitions(self, config["diff"], num_node_types, num_edge_types)

        node_dim, edge_dim = config["node_dim"], config["edge_dim"]
        time_dim = config["diff"]["time_dim"]
        self.node_embedder = nn.Linear(num_node_types, node_dim - time_dim, bias=False)
        self.edge_embedder = nn.Linear(num_edge_types, edge_dim - time_dim, bias=False)
        self.time_emb = nn.Sequential(GaussianSmearing(stop=self.num_timesteps, num_gaussians=time_dim, type_="linear"))
        den = dict(config["denoiser"])
        if den.get("backbone") != "NodeEdgeNet":
            raise NotImplementedError(den.get("backbone"))
        self.denoiser = NodeEdgeNet(node_dim, edge_dim, **den)
        self.node_decoder = MLP(node_dim, num_node_types, node_dim)
        self.edge_decoder = MLP(edge_dim, num_edge_types, edge_dim)
        self.time_dim = time_dim
        self._packed = None
        self._packed_key = None
        self.fused_transition = True     # CUDA tensors: posterior sampling in one launch (False = unfused PyTorch ops)
        self.cuda_graph = True           # sample(): replay the loop body as one CUDA graph (discrete space + fused transition)

    # ---- weights ----
    def _pack(self, device):
        return engine.PackedNet(self.state_dict(), kind=1, net_prefix="denoiser", num_blocks=self.denoiser.num_blocks,
                                update_pos=self.denoiser.update_pos, cutoff=self.denoiser.cutoff,
                                start=self.denoiser.start, time_dim=self.time_dim,
                                num_node_types=self.num_node_types, num_edge_types=self.num_edge_types,
                                num_timesteps=self.num_timesteps, device=device)

    # ---- reference API ----
    def sample_time(self, num_graphs, device, **kwargs):
        """Antithetic timestep pairs (model.py:97-104)."""
        half = torch.randint(0, self.num_timesteps, size=(num_graphs // 2 + 1,), device=device)
        time_step = torch.cat([half, self.num_timesteps - half - 1], dim=0)[:num_graphs]
        return time_step, torch.ones_like(time_step).float() / self.num_timesteps

    def _perturb(self, node_type, node_pos, batch_node, halfedge_type, batch_halfedge, time_step):
        pos_pert = self.pos_transition.add_noise(node_pos, time_step, batch_node)
        node_pert = self.node_transition.add_noise(node_type, time_step, batch_node)
        half_pert = self.edge_transition.add_noise(halfedge_type, time_step, batch_halfedge)
        return pos_pert, node_pert, half_pert

    def add_noise(self, node_type, node_pos, batch_node, halfedge_type, halfedge_index, batch_halfedge,
                  num_mol, t, bond_predictor=None, **kwargs):
        time_step = t * torch.ones(num_mol, device=node_pos.device).long()
        pos_pert, node_pert, half_pert = self._perturb(node_type, node_pos, batch_node, halfedge_type,
                                                       batch_halfedge, time_step)
        return [node_pert[0], pos_pert, half_pert[0]]

    def forward(self, h_node_pert, pos_pert, batch_node, h_edge_pert, edge_index, batch_edge, t):
        """Predict the clean molecule from its perturbed version at step t (model.py:204-234).  One C-ABI call; in a training
        step (train() mode with autograd recording) the same call is wrapped so that `loss.backward()` reaches the parameters
        (train_path.RecomputeBackward: fused forward, recompute-in-backward)."""
        plan = engine.plan_for(edge_index, h_node_pert.shape[0])

        def fused(hn, ps, he):
            return engine.moldiff_forward(self._packed_net(ps.device), plan, hn, ps, he, batch_node, batch_edge, t)

        if train_path.needs_training_backward(self):
            def recompute(hn, ps, he):
                return train_path.moldiff_forward(self, hn, ps, batch_node, he, edge_index, batch_edge, t)
            pred_node, pred_pos, pred_half = train_path.RecomputeBackward.apply(
                fused, recompute, 3, h_node_pert.float(), pos_pert.float(), h_edge_pert.float(), *self.parameters())
        else:
            pred_node, pred_pos, pred_half = fused(h_node_pert, pos_pert, h_edge_pert)
        return {"pred_node": pred_node, "pred_pos": pred_pos, "pred_halfedge": pred_half}

    def get_loss(self, node_type, node_pos, batch_node, halfedge_type, halfedge_index, batch_halfedge, num_mol):
        """Forward + diffusion losses (model.py:128-201): draw the time steps, perturb the molecule, denoise, compare."""
        time_step, _ = self.sample_time(num_mol, node_pos.device)
        pos_pert, node_pert, half_pert = self._perturb(node_type, node_pos, batch_node, halfedge_type,
                                                       batch_halfedge, time_step)
        return self.loss_from_perturbed(node_pos, batch_node, halfedge_type, halfedge_index, batch_halfedge, time_step,
                                        pos_pert, node_pert, half_pert)

    def loss_from_perturbed(self, node_pos, batch_node, halfedge_type, halfedge_index, batch_halfedge, time_step,
                            pos_pert, node_pert, half_pert):
        """The deterministic half of `get_loss` (model.py:140-201): denoise the given perturbed molecule and evaluate the
        losses.  `node_pert` / `half_pert` are what the transitions' `add_noise` returned -- (one-hot, log q(v_t), log v_0)
        in the discrete space, (v_t, v_0) in the continuous one.  Split out so that tests can teacher-force the reference's
        own random draws (tests/golden/make_golden_loss.py)."""
        edge_index = torch.cat([halfedge_index, halfedge_index.flip(0)], dim=1)
        batch_edge = torch.cat([batch_halfedge, batch_halfedge], dim=0)
        discrete = self.categorical_space == "discrete"
        if discrete:
            h_node_pert, log_node_t, log_node_0 = node_pert
            h_half_pert, log_half_t, log_half_0 = half_pert
        else:
            h_node_pert, h_node_0 = node_pert
            h_half_pert, h_half_0 = half_pert
        preds = self(h_node_pert, pos_pert, batch_node, torch.cat([h_half_pert, h_half_pert], dim=0),
                     edge_index, batch_edge, time_step)
        pred_node, pred_pos, pred_half = preds["pred_node"], preds["pred_pos"], preds["pred_halfedge"]

        out = {}
        loss_pos = F.mse_loss(pred_pos, node_pos)
        if self.bond_len_loss == True:  # noqa: E712  (config value, as in the reference)
            bond = halfedge_index[:, halfedge_type > 0]
            true_len = torch.norm(node_pos[bond[0]] - node_pos[bond[1]], dim=-1)
            pred_len = torch.norm(pred_pos[bond[0]] - pred_pos[bond[1]], dim=-1)
            out["loss_len"] = F.mse_loss(pred_len, true_len)
        if discrete:
            def vb_term(tr, logits, log_t, log_0, batch):
                log_recon = F.log_softmax(logits, dim=-1)
                post_true = tr.q_v_posterior(log_0, log_t, time_step, batch, v0_prob=True)
                post_pred = tr.q_v_posterior(log_recon, log_t, time_step, batch, v0_prob=True)
                return torch.mean(tr.compute_v_Lt(post_true, post_pred, log_0, t=time_step, batch=batch)) * 100
            loss_node = vb_term(self.node_transition, pred_node, log_node_t, log_node_0, batch_node)
            loss_edge = vb_term(self.edge_transition, pred_half, log_half_t, log_half_0, batch_halfedge)
        else:
            loss_node = F.mse_loss(pred_node, h_node_0) * 30
            loss_edge = F.mse_loss(pred_half, h_half_0) * 30
        total = loss_pos + loss_node + loss_edge + (out["loss_len"] if "loss_len" in out else 0)
        return {"loss": total, "loss_pos": loss_pos, "loss_node": loss_node, "loss_edge": loss_edge, **out}

    # ---- sampling ----
    def _guidance_delta(self, bond_predictor, gui_type, gui_scale, h_node_pert, pos_pert, batch_node,
                        edge_index, batch_edge, time_step, halfedge_type_prev, log_halfedge_type):
        """-d objective / d pos * scale (model.py:309-361).  The bond predictor's forward AND its backward
        with respect to the positions are CUDA kernels; autograd only differentiates the scalar objective
        with respect to the [Eh, K] logits."""
        with torch.enable_grad():
            pos_in = pos_pert.detach().requires_grad_(True)
            logits = bond_predictor(h_node_pert.detach(), pos_in, batch_node, edge_index, batch_edge, time_step)
            sign = -1.0
            if gui_type in ("entropy", "entropy_bond"):
                prob = torch.softmax(logits, dim=-1)
                ent = (-torch.sum(prob * torch.log(prob + 1e-12), dim=-1)).log()
                obj = ent.sum() if gui_type == "entropy" else (ent * prob[:, 1:].detach().sum(dim=-1)).sum()
            elif gui_type in ("uncertainty", "uncertainty_bond"):
                unc = torch.sigmoid(-torch.logsumexp(logits, dim=-1)).log()
                if gui_type == "uncertainty":
                    obj = unc.sum()
                else:
                    obj = (unc * torch.softmax(logits, dim=-1)[:, 1:].detach().sum(dim=-1)).sum()
            elif gui_type in ("logit_bond", "logit"):
                keep = ((halfedge_type_prev >= 1) & (halfedge_type_prev <= 4)) if gui_type == "logit_bond" \
                    else (halfedge_type_prev <= 4)
                idx = keep.nonzero().squeeze(-1)
                obj = logits[idx, halfedge_type_prev[idx]].sum()
                sign = 1.0
            elif gui_type in ("crossent", "crossent_bond"):
                if gui_type == "crossent":
                    ce = F.cross_entropy(logits, log_halfedge_type.exp()[:, :-1], reduction="none")
                else:
                    ce = F.cross_entropy(logits[:, 1:], log_halfedge_type.exp()[:, 1:-1], reduction="none")
                obj = ce.log().sum()
            else:
                raise NotImplementedError(f"Guidance type {gui_type} is not implemented")
            grad = torch.autograd.grad(obj, pos_in)[0]
        return sign * grad * gui_scale

    @torch.no_grad()
    def sample_begin(self, n_graphs, batch_node, halfedge_index, batch_halfedge):
        """Draw x_T and build the per-batch constants of the sampling loop (model.py:238-270).  Returns the
        mutable sampler state used by `sample_step` (and by bench.py, whose timed "step" is one loop body)."""
        n_nodes, n_half = len(batch_node), len(batch_halfedge)
        discrete = self.categorical_space == "discrete"
        node_init = self.node_transition.sample_init(n_nodes)
        pos = self.pos_transition.sample_init([n_nodes, 3])
        half_init = self.edge_transition.sample_init(n_half)
        st = {"n_graphs": n_graphs, "batch_node": batch_node, "batch_halfedge": batch_halfedge, "pos": pos,
              "edge_index": torch.cat([halfedge_index, halfedge_index.flip(0)], dim=1),
              "batch_edge": torch.cat([batch_halfedge, batch_halfedge], dim=0)}
        if discrete:
            _, st["h_node"], st["log_node"] = node_init
            _, st["h_half"], st["log_half"] = half_init
        else:
            st["h_node"], st["h_half"] = node_init, half_init
        return st

    @torch.no_grad()
    def sample_step(self, st, step, bond_predictor=None, guidance=None):
        """One body of the reverse loop at timestep `step` (model.py:272-372): denoise, sample x_{t-1} from the
        posteriors, add the guidance drift.  Mutates `st`; returns the step's predictions."""
        device = st["pos"].device
        discrete = self.categorical_space == "discrete"
        batch_node, batch_halfedge = st["batch_node"], st["batch_halfedge"]
        h_node, pos, h_half = st["h_node"], st["pos"], st["h_half"]
        if torch.is_tensor(step):      # a device tensor [n_graphs] (CUDA-graphed step: the value changes between replays)
            time_step = step
        else:
            time_step = torch.full((st["n_graphs"],), step, dtype=torch.long, device=device)
        h_edge = st.pop("h_edge2", None)              # [2 Eh, Ke] written directly by the fused transition step
        if h_edge is None or h_edge.data_ptr() != h_half.data_ptr():   # (stale if the caller replaced st["h_half"])
            h_edge = torch.cat([h_half, h_half], dim=0)
        preds = self(h_node, pos, batch_node, h_edge, st["edge_index"], st["batch_edge"], time_step)
        pred_node, pred_pos, pred_half = preds["pred_node"], preds["pred_pos"], preds["pred_halfedge"]

        half_type_prev = None
        fused = discrete and pos.is_cuda and self.fused_transition
        if fused:
            # one CUDA launch for the whole posterior-sampling block (csrc/mdb_transition.cuh)
            pos_prev, st["log_node"], h_node_prev, st["log_half"], h_edge2, half_type_prev = engine.transition_step(
                self.pos_transition, self.node_transition, self.edge_transition, time_step, batch_node, batch_halfedge,
                pos, pred_pos, pred_node, st["log_node"], pred_half, st["log_half"])
            h_half_prev = h_edge2[: h_half.shape[0]]
            st["h_edge2"] = h_edge2
        else:
            pos_prev = self.pos_transition.get_prev_from_recon(x_t=pos, x_recon=pred_pos, t=time_step, batch=batch_node)
        if fused:
            pass                                      # everything below was produced by the fused launch
        elif discrete:
            st["log_node"] = self.node_transition.q_v_posterior(F.log_softmax(pred_node, dim=-1), st["log_node"],
                                                                time_step, batch_node, v0_prob=True)
            h_node_prev = self.node_transition.onehot_encode(gumbel_argmax(st["log_node"]))
            st["log_half"] = self.edge_transition.q_v_posterior(F.log_softmax(pred_half, dim=-1), st["log_half"],
                                                                time_step, batch_halfedge, v0_prob=True)
            half_type_prev = gumbel_argmax(st["log_half"])
            h_half_prev = self.edge_transition.onehot_encode(half_type_prev)
        else:
            h_node_prev = self.node_transition.get_prev_from_recon(x_t=h_node, x_recon=pred_node, t=time_step,
                                                                   batch=batch_node)
            h_half_prev = self.edge_transition.get_prev_from_recon(x_t=h_half, x_recon=pred_half, t=time_step,
                                                                   batch=batch_halfedge)
        if guidance is not None:
            gui_type, gui_scale = guidance
            if gui_scale > 0:
                pos_prev = pos_prev + self._guidance_delta(
                    bond_predictor, gui_type, gui_scale, h_node, pos, batch_node, st["edge_index"], st["batch_edge"],
                    time_step, half_type_prev, st["log_half"] if discrete else None)
        st["h_node"], st["pos"], st["h_half"] = h_node_prev, pos_prev, h_half_prev
        return preds

    def graphed_step(self, st, bond_predictor=None, guidance=None):
        """CUDA-graph capture of one loop body: returns `GraphedStep`, whose `run(step)` replays the ~50
        (unguided) / ~110 (guided) kernel launches of `sample_step` as ONE graph launch -- at config-1 size (B = 32) the step
        is launch-bound.  The sampler state lives in static buffers owned by the returned object."""
        return GraphedStep(self, st, bond_predictor, guidance)

    @torch.no_grad()
    def sample(self, n_graphs, batch_node, halfedge_index, batch_halfedge, bond_predictor=None, guidance=None,
               progress=False):
        """T-step ancestral sampling (model.py:236-378).  Returns {'pred': [node, pos, halfedge] of the
        last step, 'traj': [node_traj, pos_traj, halfedge_traj]} exactly like the reference."""
        device = batch_node.device
        T = self.num_timesteps
        st = self.sample_begin(n_graphs, batch_node, halfedge_index, batch_halfedge)
        n_nodes, n_half = len(batch_node), len(batch_halfedge)
        node_traj = torch.zeros([T + 1, n_nodes, st["h_node"].shape[-1]], dtype=st["h_node"].dtype, device=device)
        pos_traj = torch.zeros([T + 1, n_nodes, 3], dtype=st["pos"].dtype, device=device)
        half_traj = torch.zeros([T + 1, n_half, st["h_half"].shape[-1]], dtype=st["h_half"].dtype, device=device)
        node_traj[0], pos_traj[0], half_traj[0] = st["h_node"], st["pos"], st["h_half"]
        steps = range(T - 1, -1, -1)
        if progress:
            from tqdm import tqdm
            steps = tqdm(steps, total=T)
        preds = None
        graphed = None
        if (getattr(self, "cuda_graph", False) and device.type == "cuda" and self.categorical_space == "discrete"
                and self.fused_transition):
            graphed = self.graphed_step(st, bond_predictor=bond_predictor, guidance=guidance)
            st = graphed.st
        for i, step in enumerate(steps):
            if graphed is not None:
                preds = graphed.run(step)
            else:
                preds = self.sample_step(st, step, bond_predictor=bond_predictor, guidance=guidance)
            if i == 0 and device.type == "cuda" and self._packed_net(device).tc_blob is not None:
                # first step: the denoiser's unbounded activations must sit inside the fp16 operand range (raises otherwise)
                engine.check_operand_range(engine.plan_for(st["edge_index"], n_nodes), self.denoiser.num_blocks)
            node_traj[i + 1], pos_traj[i + 1], half_traj[i + 1] = st["h_node"], st["pos"], st["h_half"]
        return {"pred": [preds["pred_node"], preds["pred_pos"], preds["pred_halfedge"]],
                "traj": [node_traj, pos_traj, half_traj]}


class GraphedStep:
    """One body of the reverse-diffusion loop (`MolDiff.sample_step`: denoiser, fused posterior sampling, optional bond-
    predictor guidance forward + backward) captured once as a CUDA graph and replayed per time step.  Everything the body
    reads or writes between steps is a static buffer: the sampler state (`self.st`), the time-step vector and the last
    predictions; the random variates come from torch's graph-safe Philox state.  Discrete space + fused transition only."""

    STATE = ("h_node", "pos", "log_node", "log_half")

    def __init__(self, model, st, bond_predictor=None, guidance=None, warmup=2):
        dev = st["pos"].device
        if dev.type != "cuda" or model.categorical_space != "discrete" or not model.fused_transition:
            raise engine.MoldiffB200Error("GraphedStep needs CUDA tensors, the discrete space and the fused transition step")
        self.model, self.bond, self.guidance = model, bond_predictor, guidance
        n_half = st["h_half"].shape[0]
        self.st = dict(st)
        for k in self.STATE:
            self.st[k] = st[k].clone()
        self.h_edge2 = torch.cat([st["h_half"], st["h_half"]], dim=0)      # [2 Eh, Ke]; st["h_half"] is its first half
        self.st["h_half"] = self.h_edge2[:n_half]
        self.t = torch.zeros(st["n_graphs"], dtype=torch.long, device=dev)
        saved = {k: self.st[k].clone() for k in self.STATE}
        saved_edge = self.h_edge2.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # warm-up on a side stream: lazy initialisation, plan, workspaces, packing
            self.t.fill_(model.num_timesteps - 1)
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        n0 = engine.launch_count()
        with torch.cuda.graph(self.graph):
            self.preds = self._body()
        self.launches_per_replay = engine.launch_count() - n0     # kernels of this library inside one replay
        for k in self.STATE:                          # the warm-up steps advanced the chain: rewind
            self.st[k].copy_(saved[k])
        self.h_edge2.copy_(saved_edge)

    @torch.no_grad()
    def _body(self):
        work = dict(self.st)
        work["h_edge2"] = self.h_edge2
        preds = self.model.sample_step(work, self.t, bond_predictor=self.bond, guidance=self.guidance)
        for k in self.STATE:
            self.st[k].copy_(work[k])
        self.h_edge2.copy_(work["h_edge2"])
        return preds

    def run(self, step):
        self.t.fill_(int(step))
        self.graph.replay()
        return self.preds
