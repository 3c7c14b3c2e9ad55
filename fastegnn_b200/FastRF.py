"""Host-side mirror of the reference's models/FastRF.py (the radial-field sibling of FastEGNN).

Same class names, constructor arguments, forward signature, parameter creation order (hence the same seeded
initialisation) and state_dict keys as the reference (models/FastRF.py:6-99,186-240), so
`from models.FastRF import FastRF` (main_protein.py:18,114-116; utils/train.py:57,111) keeps working.  FastRF is
a subset of the FastEGNN layer: no phi_h / phi_hv (node and virtual-node features pass through every layer
unchanged, :186) and the velocity head acts on |v_i| (Linear(1, H), :76-80,135).  It runs on the same sm_100a
kernels through fegnn_model_forward / fegnn_model_backward with FEGNN_F_RF set; there is no CPU / eager fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from .FastEGNN import FastEGNN, _coord_head, _flags, _gravity_list


class E_GCL_vel(nn.Module):
    """Parameter container of one FastRF layer (models/FastRF.py:11-89; same creation order)."""

    def __init__(self, node_feat_nf, node_feat_out_nf, node_attr_nf, edge_attr_nf, hidden_nf, virtual_channels,
                 act_fn=nn.SiLU(), residual=True, attention=False, normalize=False, coords_agg='mean', tanh=False,
                 gravity=None):
        super().__init__()
        if hidden_nf != L.H or node_feat_nf != L.H or node_feat_out_nf != L.H:
            raise NotImplementedError(f"the sm_100a kernels are built for hidden_nf == {L.H}")
        if not isinstance(act_fn, nn.SiLU):
            raise NotImplementedError("only act_fn=nn.SiLU() is implemented (the reference default)")
        if coords_agg != 'mean':
            raise Exception('Wrong coords_agg parameter')      # models/FastRF.py:116
        if not 1 <= virtual_channels <= L.MAX_C:
            raise NotImplementedError(f"virtual_channels must be in [1, {L.MAX_C}]")
        if not 0 <= edge_attr_nf <= L.MAX_FE:
            raise NotImplementedError(f"edge_attr_nf must be in [0, {L.MAX_FE}]")
        self.residual, self.attention, self.normalize, self.coords_agg, self.tanh = \
            residual, attention, normalize, coords_agg, tanh
        self.hiddden_nf = hidden_nf            # sic (:18)
        self.node_feat_out_nf = node_feat_out_nf
        self.epsilon = 1e-8
        self.virtual_channels = virtual_channels
        H, Cc = hidden_nf, virtual_channels
        self.edge_mlp = nn.Sequential(nn.Linear(2 * H + 1 + edge_attr_nf, H), act_fn, nn.Linear(H, H), act_fn)
        self.edge_mlp_virtual = nn.Sequential(nn.Linear(2 * H + 1 + Cc, H), act_fn, nn.Linear(H, H), act_fn)
        if attention:
            self.att_mlp = nn.Sequential(nn.Linear(H, 1), nn.Sigmoid())
            self.att_mlp_virtual = nn.Sequential(nn.Linear(H, 1), nn.Sigmoid())
        self.coord_mlp_r = _coord_head(H, act_fn, tanh)
        self.coord_mlp_r_virtual = _coord_head(H, act_fn, tanh)
        self.coord_mlp_v_virtual = _coord_head(H, act_fn, tanh)
        self.coord_mlp_vel = nn.Sequential(nn.Linear(1, H), act_fn, nn.Linear(H, 1))      # acts on |v| (:76-80)
        self.gravity = gravity
        if gravity is not None:
            self.gravity_mlp = nn.Sequential(nn.Linear(H, H), act_fn, nn.Linear(H, 1))

    def forward(self, *args, **kwargs):
        raise NotImplementedError("the FastRF layer is driven by FastRF.forward (fegnn_model_forward with FEGNN_F_RF); "
                                  "nothing in the reference calls the layer on its own")


class FastRF(FastEGNN):
    """Drop-in for the reference FastRF (models/FastRF.py:189-240).  Shares FastEGNN's pointer tables and autograd
    function; only the layer container and the flag word differ."""

    def __init__(self, node_feat_nf, node_attr_nf, edge_attr_nf, hidden_nf, virtual_channels, device='cpu',
                 act_fn=nn.SiLU(), n_layers=4, residual=True, attention=False, normalize=False, tanh=False,
                 gravity=None):
        nn.Module.__init__(self)
        self.hidden_nf = hidden_nf
        self.device = device
        self.n_layers = n_layers
        self.virtual_channels = virtual_channels
        assert virtual_channels > 0, f'Channels of virtual node must greater than 0 (got {virtual_channels})'
        if not 1 <= n_layers <= 32:
            raise NotImplementedError("n_layers must be in [1, 32]")
        if not 1 <= node_feat_nf <= 16:
            raise NotImplementedError("node_feat_nf must be in [1, 16]")
        if node_attr_nf != 0:
            raise NotImplementedError("node_attr_nf must be 0 (main_protein.py:115 passes 0)")
        self.virtual_node_feat = nn.Parameter(data=torch.randn(size=(1, hidden_nf, virtual_channels)),
                                              requires_grad=True)
        self.embedding_in = nn.Linear(node_feat_nf, hidden_nf)
        self._gravity = _gravity_list(gravity)
        if gravity is not None:
            gravity = torch.tensor(gravity, device=device)
        for i in range(n_layers):
            self.add_module("gcl_%d" % i, E_GCL_vel(hidden_nf, hidden_nf, node_attr_nf, edge_attr_nf, hidden_nf,
                                                    virtual_channels=virtual_channels, act_fn=act_fn,
                                                    residual=residual, attention=attention, normalize=normalize,
                                                    tanh=tanh, gravity=gravity))
        self._flag_word = _flags(attention, normalize, tanh, gravity) | L.F_RF
        self._edge_attr_nf = edge_attr_nf
        self._cache = None
        self._dead_names = frozenset()          # every FastRF parameter receives a gradient
        self.to(self.device)

    def _signature_tensors(self):
        return (self.virtual_node_feat, self.embedding_in.weight,
                getattr(self, "gcl_%d" % (self.n_layers - 1)).coord_mlp_vel[2].bias)
