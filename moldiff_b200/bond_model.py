"""``BondPredictor`` -- the guidance network (reference ``models/bond_predictor.py:11-162``), B200-native.

Same constructor / ``forward`` / ``get_loss`` surface and the same 554-key ``state_dict``.  ``forward``
is a ``torch.autograd.Function`` whose forward is ``mdb_bondpred_forward`` (8 NodeEdgeNet blocks with
``update_pos=False`` + the 3-layer edge decoder, all sm_100a kernels) and whose backward with respect
to ``pos_node`` is ``mdb_bondpred_backward`` (hand-written input-gradient kernels that recompute the
per-edge activations tile by tile instead of saving ~275 KB/edge like eager autograd).  That is the
only gradient the sampling guidance needs (``models/model.py:312-325``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import engine, train_path
from .diffusion_model import _PackedMixin, build_transitions
from .nets import MLP, GaussianSmearing, NodeEdgeNet


class _BondLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, module, h_node, batch_node, edge_index, batch_edge, t):
        plan = engine.plan_for(edge_index, h_node.shape[0])
        net = module._packed_net(pos.device)
        need_grad = pos.requires_grad
        logits = engine.bondpred_forward(net, plan, h_node, pos, batch_node, batch_edge, t, save=need_grad)
        ctx.net, ctx.plan, ctx.generation = net, plan, plan.save_generation
        ctx.save_for_backward(h_node, pos, batch_node, batch_edge, t)
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        h_node, pos, batch_node, batch_edge, t = ctx.saved_tensors
        d_pos = engine.bondpred_backward(ctx.net, ctx.plan, h_node, pos, batch_node, batch_edge, t, grad_logits,
                                         generation=ctx.generation)
        return d_pos, None, None, None, None, None, None


class BondPredictor(nn.Module, _PackedMixin):
    def __init__(self, config, num_node_types, num_edge_types, **kwargs):
        super().__init__()
        self.config = config
        self.num_node_types = num_node_types
        self.num_edge_types = num_edge_types
        self.num_timesteps = config["diff"]["num_timesteps"]
        if self.num_timesteps != 0:
            build_transitions(self, config["diff"], num_node_types, num_edge_types, with_edges=False)
        node_dim, edge_dim = config["node_dim"], config["edge_dim"]
        time_dim = config["diff"]["time_dim"] if self.num_timesteps > 0 else 0
        self.node_embedder = nn.Linear(num_node_types, node_dim - time_dim, bias=False)
        self.edge_embedder = nn.Linear(num_node_types * 2, edge_dim - time_dim, bias=False)
        if self.num_timesteps != 0:
            self.time_emb = GaussianSmearing(stop=self.num_timesteps, num_gaussians=time_dim, type_="linear")
        self.encoder = NodeEdgeNet(node_dim, edge_dim, **dict(config["encoder"]))
        if self.encoder.update_pos:
            raise NotImplementedError("BondPredictor kernels assume encoder.update_pos = False (train_bondpred.yml)")
        self.edge_decoder = MLP(edge_dim + node_dim, num_edge_types, edge_dim, num_layer=3)
        self.edge_weight = torch.tensor([0.1] + [1.0] * (self.num_edge_types - 1), dtype=torch.float32)
        self.ce_loss = torch.nn.CrossEntropyLoss(self.edge_weight)
        self.time_dim = time_dim
        self._packed = None
        self._packed_key = None

    def _pack(self, device):
        return engine.PackedNet(self.state_dict(), kind=2, net_prefix="encoder", num_blocks=self.encoder.num_blocks,
                                update_pos=False, cutoff=self.encoder.cutoff, start=self.encoder.start,
                                time_dim=self.time_dim, num_node_types=self.num_node_types,
                                num_edge_types=self.num_edge_types, num_timesteps=max(self.num_timesteps, 1),
                                device=device)

    def sample_time(self, num_graphs, device, **kwargs):
        half = torch.randint(0, self.num_timesteps, size=(num_graphs // 2 + 1,), device=device)
        time_step = torch.cat([half, self.num_timesteps - half - 1], dim=0)[:num_graphs]
        return time_step, torch.ones_like(time_step).float() / self.num_timesteps

    def forward(self, h_node, pos_node, batch_node, edge_index, batch_edge, t):
        """Bond-type logits for the half edges, [E/2, num_edge_types] (bond_predictor.py:128-162)."""
        if self.num_timesteps == 0:
            t = torch.zeros(int(batch_node.max()) + 1, device=pos_node.device, dtype=torch.long)
        if train_path.needs_training_backward(self):          # train_bond.py: weight gradients (fused forward, recompute-in-backward)
            plan = engine.plan_for(edge_index, h_node.shape[0])

            def fused(hn, ps):
                return engine.bondpred_forward(self._packed_net(ps.device), plan, hn, ps, batch_node, batch_edge, t)

            def recompute(hn, ps):
                return train_path.bondpred_forward(self, hn, ps, batch_node, edge_index, batch_edge, t)
            return train_path.RecomputeBackward.apply(fused, recompute, 2, h_node.float(), pos_node.float(), *self.parameters())
        return _BondLogits.apply(pos_node, self, h_node, batch_node, edge_index, batch_edge, t)

    def get_loss(self, node_type, node_pos, batch_node, halfedge_type, halfedge_index, batch_halfedge, num_mol):
        """Weighted cross-entropy of the predicted half-edge types (bond_predictor.py:84-124)."""
        device = node_pos.device
        if self.num_timesteps != 0:
            time_step, _ = self.sample_time(num_mol, device)
            pos_node = self.pos_transition.add_noise(node_pos, time_step, batch_node)
            h_node = self.node_transition.add_noise(node_type, time_step, batch_node)[0]
        else:
            time_step = None
            h_node = F.one_hot(node_type, self.num_node_types).float()
            pos_node = node_pos
        edge_index = torch.cat([halfedge_index, halfedge_index.flip(0)], dim=1)
        batch_edge = torch.cat([batch_halfedge, batch_halfedge], dim=0)
        pred = self(h_node, pos_node, batch_node, edge_index, batch_edge, time_step)
        loss_edge = self.ce_loss(pred, halfedge_type)
        return {"loss": loss_edge, "loss_edge": loss_edge}
