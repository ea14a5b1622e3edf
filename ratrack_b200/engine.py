"""Host side of the fused inference engine (include/ratrack_b200.h, section 2).

`FusedBackbone(module)` reads the parameters of a `Track4DBackbone` (state_dict surface of the reference's
Track4D: pn_head.*, fc_layer.*, fd_layer.*), folds every eval-mode BatchNorm into the 1x1 convolution in
front of it, splits first-layer weights into their xyz / feature / cloud-constant column blocks, and hands
the resulting device tensors to the C-ABI engine as one pointer table.  Calling it runs
`rt_backbone_forward` on torch's current stream and returns the 7-tuple of `Track4D.backbone`
(reference: src/models/track4d.py:67-86).  There is no fallback: a missing library or a failed launch raises.

Folding (reference layer -> engine tensor):
  SharedMLP layer  conv(no bias) -> BN(eval, eps 1e-5) -> ReLU   (src/lib/pytorch_utils.py:5-32)
      W' = W * gamma / sqrt(var + eps),  b' = beta - mean * gamma / sqrt(var + eps)
  first SA conv (C1 x (3 + Cin)):  WX = W'[:, :3]  (applied to xyz_j - centre),  WF = W'[:, 3:]  (applied per point)
  cost-volume conv 0 (256 x 515):  columns [f1 local 128 | f1 global 128 | f2 local 128 | f2 global 128 | dxyz 3]
"""
import ctypes

import torch

from . import _cabi


def _fold(conv_w, bn):
    """conv weight (Cout,Cin,1,1) + BatchNorm2d -> (W' (Cout,Cin), b' (Cout,)) in float64 then fp32."""
    w = conv_w.detach().double().flatten(1)
    scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    return (w * scale[:, None]).float(), (bn.bias.detach().double() - bn.running_mean.detach().double() * scale).float()


def _shared_mlp_layers(mlp):
    """SharedMLP -> list of folded (W', b')."""
    out = []
    for layer in mlp:
        out.append(_fold(layer.conv.weight, layer.bn.bn))
    return out


def _head_weights(head, first_feature_split):
    """Weight list of one PNHead in the order of struct HeadW (csrc/engine.cu)."""
    ws = []
    for lvl, sa in enumerate((head.sa1, head.sa2, head.sa3)):
        scales = [_shared_mlp_layers(m) for m in sa.mlps]
        wf = torch.cat([s[0][0][:, 3:] for s in scales], dim=0)  # feature columns of both scales' first conv
        if lvl == 0:
            cols = 0
            for width in first_feature_split:
                ws.append(wf[:, cols:cols + width] if width else None)
                cols += width
            assert cols == wf.shape[1], (cols, wf.shape)
        else:
            ws.append(wf)
        for s in scales:
            ws += [s[0][0][:, :3], s[0][1], s[1][0], s[1][1]]
            ws += [s[2][0], s[2][1]] if len(s) > 2 else [None, None]
        lin = (head.linear1, head.linear2, head.linear3)[lvl]
        ws += [lin.weight.detach(), lin.bias.detach()]
    w3, b3 = _shared_mlp_layers(head.fp3.mlp)[0]
    w2, b2 = _shared_mlp_layers(head.fp2.mlp)[0]
    w1, b1 = _shared_mlp_layers(head.fp1.mlp)[0]
    ws += [w3[:, :64], w3[:, 64:], b3, w2[:, :128], w2[:, 128:], b2, w1, b1]
    assert len(ws) == 56, len(ws)
    return ws


def _weightnet(wn):
    c = wn.mlp_convs
    assert not wn.bn and len(c) == 3
    out = []
    for conv in c:
        out += [conv.weight.detach().flatten(1), conv.bias.detach()]
    return out


def _predictor(p):
    """FlowPredictor / ClsPredictor trunk: three folded conv+BN layers and the final bias-free conv."""
    out = []
    for block in p.sf_mlp:
        out.append(_fold(block[0].weight, block[1]))
    return out, p.conv2.weight.detach().flatten(1)


LO_SCALE = 2048.0
W_SCALE = 1024.0  # costvol_tc.cu: weights enter the tensor core as 2^10 * W (keeps the fp16 `lo` plane normal)


def _umma_planes(w, k_chunk):
    """(N, K) fp32 -> fp16 hi/lo planes in the K-major core-matrix layout of tcgen05 shared-memory operands:
    [chunk = K / k_chunk][plane hi, lo][kc = k_chunk / 8][row group = N / 8][8 rows][8 halfs]."""
    n, k = w.shape
    assert n % 8 == 0 and k % k_chunk == 0 and k_chunk % 8 == 0
    w = w.float() * W_SCALE
    hi = w.half()
    lo = ((w - hi.float()) * LO_SCALE).half()     # lo plane stored as 2^11 * lo (costvol_tc.cu / mlp_tc.cu: scale-input-d)

    def lay(x):
        return x.reshape(n // 8, 8, k // k_chunk, k_chunk // 8, 8).permute(2, 3, 0, 1, 4)

    return torch.stack([lay(hi), lay(lo)], dim=1).contiguous()


def pack_costvol_weights(w2, w3):
    """conv 256->256 weights of cost-volume layers 2 and 3 -> one fp16 tensor streamed chunk by chunk (K = 32)."""
    return torch.cat([_umma_planes(w2, 32), _umma_planes(w3, 32)], dim=0)


def pack_weightnet_last(wc):
    """WeightNet last conv (256 x 8) -> K padded to 16, planes [hi, lo][kc 2][32][8][8]."""
    wc16 = torch.cat([wc.float(), torch.zeros(wc.shape[0], 8, dtype=torch.float32, device=wc.device)], dim=1)
    return _umma_planes(wc16, 16)[0]


def build_weight_table(net):
    """-> list of 157 tensors / None, ordered as struct EngineW (fp32, except the two fp16 packs of CvW)."""
    fc, fd = net.fc_layer, net.fd_layer
    assert not fc.bn and fc.nsample == 16, "engine is built for the reference configuration (bn=False, nsample=16)"
    ws = _head_weights(net.pn_head, (2, 0, 0, 0))
    ws += _head_weights(fd.mse, (2, 128, 128, 256))
    w1 = fc.mlp_convs[0].weight.detach().flatten(1)
    assert w1.shape == (256, 515)
    ws += [w1[:, 0:128], w1[:, 128:256], w1[:, 256:384], w1[:, 384:512], w1[:, 512:515], fc.mlp_convs[0].bias.detach(),
           fc.mlp_convs[1].weight.detach().flatten(1), fc.mlp_convs[1].bias.detach(),
           fc.mlp_convs[2].weight.detach().flatten(1), fc.mlp_convs[2].bias.detach()]
    ws += _weightnet(fc.weightnet1) + _weightnet(fc.weightnet2)
    ws += [pack_costvol_weights(fc.mlp_convs[1].weight.detach().flatten(1), fc.mlp_convs[2].weight.detach().flatten(1)),
           pack_weightnet_last(fc.weightnet1.mlp_convs[2].weight.detach().flatten(1))]
    (c1, c2, c3), c4 = _predictor(fd.cp)
    ws += [c1[0], c1[1], c2[0], c2[1], c3[0], c3[1], c4, fd.cp.linear.weight.detach(), fd.cp.linear.bias.detach()]
    (f1, f2, f3), f4 = _predictor(fd.fp)
    ws += [f1[0][:, :128], f1[0][:, 128:], f1[1], f2[0], f2[1], f3[0], f3[1], f4]
    g = fd.torchGRU
    assert g.num_layers == 5 and g.input_size == 128 and g.hidden_size == 128
    # transposed to (5, 128, 384) so the GRU kernel's per-gate-row threads read coalesced columns
    ws += [torch.stack([getattr(g, f"weight_ih_l{l}").detach().t() for l in range(5)]),
           torch.stack([getattr(g, f"weight_hh_l{l}").detach().t() for l in range(5)]),
           torch.stack([getattr(g, f"bias_ih_l{l}").detach() for l in range(5)]),
           torch.stack([getattr(g, f"bias_hh_l{l}").detach() for l in range(5)])]
    return ws


class FusedBackbone:
    def __init__(self, net):
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise _cabi.RatrackError("FusedBackbone needs the module on a CUDA device (there is no CPU path)")
        self.device = dev
        self.npoint = int(net.npoints_cfg)
        table = build_weight_table(net)
        lib = _cabi.lib()
        assert len(table) == lib.rt_engine_num_weights(), (len(table), lib.rt_engine_num_weights())
        # keep the folded tensors alive: the engine stores raw pointers
        self._weights = [None if t is None else
                         t.to(device=dev, dtype=torch.float16 if t.dtype == torch.float16 else torch.float32).contiguous().clone()
                         for t in table]
        ptrs = (ctypes.c_void_p * len(table))(*[None if t is None else t.data_ptr() for t in self._weights])
        self._handle = ctypes.c_void_p()
        _cabi.call("rt_engine_create", ctypes.byref(self._handle), self.npoint, ptrs, len(table))
        self._workspace = None
        self._ws_key = None
        self.last_knn = None
        self._prof = None
        # fp16-range guard of the tensor-core kernels (costvol_tc.cu / mlp_tc.cu): weights enter as 2^10 * W, activations
        # as fp16 hi/lo planes.  Out-of-range WEIGHTS are known now: such a model runs on the fp32 SIMT kernels.
        # Out-of-range ACTIVATIONS are caught per forward: the device overwrites that step's outputs with NaN, the status
        # word comes back without a host synchronisation and the next forward (and everything after it) uses the fp32
        # SIMT kernels -- see _poll_status / run_checked.
        self.tensor_cores = True
        self._status_host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self._status_event = None
        wmax = max(float(t.abs().max()) for t in self._weights if t is not None and t.dtype == torch.float32 and t.dim() >= 2)   # matrices: what gets packed
        if not wmax * W_SCALE < 65000.0:
            self._to_simt(f"a folded weight has magnitude {wmax:.3g}: 2^10 * W leaves the fp16 range")

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _cabi.lib().rt_engine_destroy(h)
            except Exception:
                pass

    def set_flags(self, costvol_tc: bool = True, mlp_tc: bool = True, two_lanes: bool = False, fps_exclusive: bool = True,
                  knn_early: bool = True, own_stream: bool = True, morton: bool = True, fps_identity: bool = True):
        """A/B switches.  Arithmetic: tcgen05 kernels (default) vs the fp32 SIMT kernels of the same dataflow,
        separately for the cost-volume core (costvol_tc.cu) and for every other dense layer (mlp_tc.cu).
        Scheduling (results are bit-identical either way): two_lanes = run the two halves of a batch (>= 8 pairs)
        concurrently on separate stream sets; fps_exclusive = every FPS CTA claims a whole SM so co-running kernels
        cannot stretch its latency chain; knn_early = the cost-volume kNN runs beside the FPS chain instead of after
        it; own_stream = the feature path runs on an engine-owned stream of middle priority; morton = the cost-volume
        kernels walk pc1 in Morton order so that the points of a tile share gathered neighbour rows; fps_identity = levels 2
        and 3 of the FPS chain (512 of 512 points) are answered by a parallel identity check where it proves the serial result."""
        _cabi.call("rt_engine_set_flags", self._handle,
                   (1 if costvol_tc else 0) | (2 if mlp_tc else 0) | (4 if two_lanes else 0) | (8 if fps_exclusive else 0) |
                   (16 if knn_early else 0) | (32 if own_stream else 0) | (64 if morton else 0) | (512 if fps_identity else 0))
        self._ws_key = None   # the workspace layout depends on the lane split

    def _to_simt(self, why):
        import warnings

        warnings.warn(f"ratrack_b200: {why}; this engine now runs its dense layers on the fp32 SIMT kernels "
                      "(same dataflow, ~3x slower)", RuntimeWarning, stacklevel=3)
        self.tensor_cores = False
        self.set_flags(costvol_tc=False, mlp_tc=False)

    def _poll_status(self, block=False):
        """Status of the previous forward, if it has arrived (block=True waits for it).  True = its guard fired (its
        outputs are NaN on the device) and the engine has switched to the fp32 SIMT kernels."""
        ev = self._status_event
        if ev is None or not (block or ev.query()):
            return False
        if block:
            ev.synchronize()
        self._status_event = None
        if int(self._status_host[0]) | int(self._status_host[1]):
            if self.tensor_cores:
                self._to_simt("an activation left the fp16 hi/lo range of the tensor-core kernels (that forward's outputs were "
                              "overwritten with NaN)")
            return True
        return False

    def run_checked(self, *args, **kw):
        """Forward + blocking status check; a forward whose fp16-range guard fired is repeated on the fp32 SIMT kernels."""
        out = self(*args, **kw)
        if self._poll_status(block=True):
            out = self(*args, **kw)
            if self._poll_status(block=True):
                raise _cabi.RatrackError("fused backbone: device status set on the fp32 SIMT path")
        return out

    def num_lanes(self, batch):
        return int(_cabi.lib().rt_engine_num_lanes(self._handle, int(batch)))

    def check_status(self):
        """Blocking.  Raises if the last forward flagged an fp16-range overflow in the tensor-core kernels."""
        st = ctypes.c_int(0)
        _cabi.call("rt_engine_last_status", self._handle, ctypes.byref(st))
        if st.value:
            raise _cabi.RatrackError(f"fused backbone: device status {st.value} (cost-volume activation outside fp16 hi/lo range)")

    def stage_profile(self, on=True):
        """Record CUDA events at the stage boundaries of the next forwards (feature-path stream)."""
        _cabi.call("rt_engine_stage_profile", self._handle, 1 if on else 0)

    def stage_times(self):
        """[(stage name, milliseconds)] of the last forward (blocking)."""
        ms = (ctypes.c_float * 16)()
        names = (ctypes.c_char_p * 16)()
        n = _cabi.lib().rt_engine_stage_times(self._handle, ms, names, 16)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def launch_count(self):
        return int(_cabi.lib().rt_engine_launch_count(self._handle))

    def set_profile_events(self, start, stop):
        """torch.cuda.Event pair (enable_timing=True) recorded around the dominant kernel of every forward."""
        if start is None:
            _cabi.call("rt_engine_set_profile_events", self._handle, None, None)
            self._prof = None
            return
        # events must exist on the device before their handle can be taken
        start.record()
        stop.record()
        self._prof = (start, stop)
        _cabi.call("rt_engine_set_profile_events", self._handle, start.cuda_event, stop.cuda_event)

    def _ws(self, b, n):
        if self._ws_key != (b, n):
            nbytes = int(_cabi.lib().rt_engine_workspace_bytes(self._handle, b, n))
            if nbytes < 0:
                raise _cabi.RatrackError("rt_engine_workspace_bytes failed")
            self._workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._ws_key = (b, n)
        base = self._workspace.data_ptr()
        return (base + 255) // 256 * 256, self._workspace.numel() - 256

    def __call__(self, pc1, pc2, feature1, feature2, h=None, want_knn=False, npts1=None, npts2=None):
        """npts1 / npts2: (B,) point counts of a zero-padded variable-size batch (data_io.PaddedBatcher); every pair is then
        computed exactly as it would be alone at its own size, output columns of padded points are zero."""
        self._poll_status()
        b, _, n = pc1.shape
        if (npts1 is None) != (npts2 is None):
            raise _cabi.RatrackError("fused backbone: npts1 and npts2 go together")
        if npts1 is not None:
            counts = [torch.as_tensor(x, dtype=torch.int32).cpu().contiguous() for x in (npts1, npts2)]
            if any(c.numel() != b for c in counts):
                raise _cabi.RatrackError(f"fused backbone: point counts must have one entry per pair ({b})")
        f32 = dict(dtype=torch.float32, device=self.device)
        args = [t.to(**f32).contiguous() for t in (pc1, pc2, feature1, feature2)]
        if h is None:
            h = torch.zeros(5, b, 128, **f32)
        h = h.to(**f32).contiguous()
        flow = torch.empty(b, 3, n, **f32)
        h_out = torch.empty(5, b, 128, **f32)
        cls = torch.empty(b, n, **f32)
        cor = torch.empty(b, 256, n, **f32)
        f1 = torch.empty(b, 256, n, **f32)
        f2 = torch.empty(b, 256, n, **f32)
        prop = torch.empty(b, 128, n, **f32)
        knn = (torch.empty(b, n, 16, dtype=torch.int32, device=self.device),
               torch.empty(b, n, 16, dtype=torch.int32, device=self.device)) if want_knn else (None, None)
        ws_ptr, ws_bytes = self._ws(b, n)
        with torch.cuda.device(self.device):
            tail = (args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(), h.data_ptr(), flow.data_ptr(),
                    h_out.data_ptr(), cls.data_ptr(), cor.data_ptr(), f1.data_ptr(), f2.data_ptr(), prop.data_ptr(),
                    knn[0].data_ptr() if want_knn else None, knn[1].data_ptr() if want_knn else None,
                    ws_ptr, ws_bytes, torch.cuda.current_stream(self.device).cuda_stream)
            if npts1 is None:
                _cabi.call("rt_backbone_forward", self._handle, b, n, *tail)
            else:
                _cabi.call("rt_backbone_forward_varlen", self._handle, b, n, counts[0].data_ptr(), counts[1].data_ptr(), *tail)
            if self.tensor_cores and self._status_event is None:
                _cabi.call("rt_engine_status_async", self._handle, self._status_host.data_ptr(),
                           torch.cuda.current_stream(self.device).cuda_stream)
                self._status_event = torch.cuda.Event()
                self._status_event.record()
        if want_knn:
            self.last_knn = knn
        return flow, h_out, cls, cor, f1, f2, prop


def roofline_of_dominant(batch, points, dom_ms, peaks, launch_pairs=None):
    """Roofline entry of bench.py for the dominant kernel: the dense cost-volume MLP
    (two 256 -> 256 layers over batch*points*16 rows; DESIGN.md "Kernels and rooflines").
    Algorithmic work = the fp32-accurate useful FLOPs, 2 layers x 2*rows*256*256; the denominator is the
    measured dense bf16 tensor throughput (sustained figure: the kernel is timed inside a long step)."""
    # the timed launch is lane 0's kernel: it covers launch_pairs of the batch's pairs (all of them with one lane)
    rows = (launch_pairs or batch) * points * 16
    flops = 2 * 2.0 * rows * 256 * 256
    avg_ms = sum(dom_ms) / max(1, len(dom_ms))
    ach = flops / (avg_ms * 1e-3) / 1e12
    peak = peaks.get("bf16_sustained") or peaks["bf16"]
    traffic = None   # dram bytes of one launch from the committed ncu --set full capture, when it is this configuration
    try:
        import json
        import os
        t = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                        "costvol_traffic.json")))
        if t["batch"] == (launch_pairs or batch) and t["points"] == points:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    return {"kernel": "costvol_tc_kernel: gather + cost-volume MLP (2 x [rows x 256 x 256], LeakyReLU) + WeightNet sum",
            "bound": "tensor", "achieved": ach, "peak": peak, "peak_source": peaks["src"] + " dense bf16 (sustained)",
            "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "avg_launch_ms": avg_ms, "flops_per_launch": flops,
            "traffic_source": "profiles/costvol_traffic.json: dram__bytes_read + dram__bytes_write of one launch of this kernel (ncu --set full capture, tools/collect_profiles.sh)",
            "note": "achieved counts the fp32-accurate useful FLOPs; the tensor pipe executes 3x that (fp16 hi/lo split products)"}
