"""Per-step weight re-packing for the training step as ONE kernel launch (dn_pack_weights).

``diffnorm_b200.packing`` builds the bf16 K-major tiles with torch indexing: right for a load-time job, but the training
step has to redo it after every optimizer update, and ~500 small indexing kernels cost 10.8 ms of a 32 ms step.  Here the
same layouts are written as a table of copy descriptors (``dn_pack_op``, include/diffnorm_b200.h): the trainer packs once
with ``packing`` (which also zero-fills the padding), records next to every ``pack_*`` call which parameter feeds which
packed tensor, and from then on refreshes the packed tensors in place with one launch.

Every recorder method mirrors one ``packing`` function (same argument meaning); ``tests/test_repack_cpu.py`` runs the table
through a plain-torch executor and compares with ``packing`` bit for bit, ``tests/test_train_gpu.py`` does the same with the
CUDA kernel on the full model.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

from . import _lib

TILE = 64   # = PK_TILE in csrc/pack.cu
BIG = 1 << 30   # "no blocking": r / BIG == 0


def plan_tensors(obj, prefix: str = ""):
    """(name, tensor) for every tensor reachable from a tree of GemmPlans / lists / plain holder objects."""
    from .ops import GemmPlan
    out = []
    if isinstance(obj, GemmPlan):
        for k in ("W", "bias", "bias2"):
            t = getattr(obj, k)
            if t is not None:
                out.append((f"{prefix}{obj.name}.{k}", t))
    elif isinstance(obj, torch.Tensor):
        out.append((prefix, obj))
    elif isinstance(obj, (list, tuple)):
        for i, o in enumerate(obj):
            out += plan_tensors(o, f"{prefix}[{i}]")
    elif hasattr(obj, "__dict__"):
        for k, v in vars(obj).items():
            out += plan_tensors(v, f"{prefix}{k}:")
    return out


class PackTable:
    def __init__(self):
        self.ops: List[dict] = []
        self.extra: List[Callable[[], None]] = []   # the few packings that need arithmetic (sum of the skip biases)
        self._dev_ops = None
        self._keep = []                             # tensors whose storage the table points into
        self.total_tiles = 0

    # ------------------------------------------------------------------------------------------------ generic op
    def add(self, src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, s_row: int, s_col: int, *, taps: int = 1,
            s_tap: int = 0, src_off: int = 0, ldd: Optional[int] = None, row0: int = 0, rblk: int = BIG, rblk_stride: int = 0,
            col0: int = 0, cblk: int = BIG, cblk_stride: int = 0, tap_cols: int = 0, tap_pos: Sequence[int] = (0, 1, 2),
            r_fastest: bool = False):
        assert src.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
        assert dst.dtype in (torch.bfloat16, torch.float32) and 1 <= taps <= 3
        if rows <= 0 or cols <= 0:
            return
        self._keep += [src, dst]
        self.ops.append(dict(src=src, dst=dst, src_off=src_off, rows=rows, cols=cols, taps=taps, s_row=s_row, s_col=s_col,
                             s_tap=s_tap, ldd=dst.shape[-1] if ldd is None else ldd, row0=row0, rblk=rblk,
                             rblk_stride=rblk_stride, col0=col0, cblk=cblk, cblk_stride=cblk_stride, tap_cols=tap_cols,
                             tap_pos=tuple(tap_pos), r_fastest=bool(r_fastest), out_f32=dst.dtype == torch.float32))

    # ------------------------------------------------------------------------------------------------ packing.* mirrors
    def linear(self, W: torch.Tensor, dst: torch.Tensor, *, transposed: bool = False, row0: int = 0, col0: int = 0):
        """pack_linear(W) / pack_linear(W.t()): W [N, K] (trailing 1s squeezed) -> dst[row0 + n, col0 + k] (or [k, n])."""
        N, K = W.shape[0], W.numel() // W.shape[0]
        if transposed:
            self.add(W, dst, K, N, 1, K, row0=row0, col0=col0, r_fastest=True)
        else:
            self.add(W, dst, N, K, K, 1, row0=row0, col0=col0)

    def vector(self, b: torch.Tensor, dst: torch.Tensor, *, row0: int = 0, rblk: int = BIG, rblk_stride: int = 0):
        """A bias: dst[row0 + blockmap(n)] = b[n] (fp32)."""
        self.add(b, dst, b.numel(), 1, 1, 0, ldd=1, row0=row0, rblk=rblk, rblk_stride=rblk_stride)

    def conv3(self, W: torch.Tensor, dst: torch.Tensor, c_pad: int, *, transposed: bool = False):
        """pack_conv3(W, cin_pad=c_pad) with W [N, Cin, 3]: dst[n, k * c_pad + cin];  transposed = pack_conv3(W.permute(1, 0,
        2), cin_pad=c_pad): dst[cin, k * c_pad + n]."""
        N, Cin, _ = W.shape
        if transposed:
            self.add(W, dst, Cin, N, 3, 3 * Cin, taps=3, s_tap=1, tap_cols=c_pad, r_fastest=True)
        else:
            self.add(W, dst, N, Cin, 3 * Cin, 3, taps=3, s_tap=1, tap_cols=c_pad)

    def geglu(self, W: torch.Tensor, bias: torch.Tensor, Wp: torch.Tensor, bp: torch.Tensor, WpT: Optional[torch.Tensor] = None):
        """pack_geglu: x row n -> packed row (n / 128) * 256 + n % 128, gate row n -> + 128; WpT = Wp.t().contiguous()."""
        inner, K = W.shape[0] // 2, W.shape[1]
        for half in (0, 1):
            self.add(W, Wp, inner, K, K, 1, src_off=half * inner * K, row0=128 * half, rblk=128, rblk_stride=256)
            self.add(bias, bp, inner, 1, 1, 0, src_off=half * inner, ldd=1, row0=128 * half, rblk=128, rblk_stride=256)
            if WpT is not None:
                self.add(W, WpT, K, inner, 1, K, src_off=half * inner * K, col0=128 * half, cblk=128, cblk_stride=256,
                         r_fastest=True)

    def wavenet_level(self, convs, conv_b, ress, res_b, Wp: torch.Tensor, c_pad: int, bi: Optional[torch.Tensor] = None,
                      bc: Optional[torch.Tensor] = None, br: Optional[torch.Tensor] = None):
        """pack_wavenet_level: chain g, out channel n -> conv row g * rows_g + (n / 128) * 256 + n % 128 with K layout
        [tap2 | tap0 | tap1], res row + 128 (K position 0).  bi = the un-fused form's bias [G, tiles, (conv | res), 128]."""
        Cc = convs[0].shape[0]
        rows_g = (c_pad // 128) * 256
        for g in range(len(convs)):
            self.add(convs[g], Wp, Cc, Cc, 3 * Cc, 3, taps=3, s_tap=1, row0=g * rows_g, rblk=128, rblk_stride=256,
                     tap_cols=c_pad, tap_pos=(1, 2, 0))
            self.add(ress[g], Wp, Cc, Cc, Cc, 1, row0=g * rows_g + 128, rblk=128, rblk_stride=256)
            if bi is not None:
                self.add(conv_b[g], bi, Cc, 1, 1, 0, ldd=1, row0=g * 2 * c_pad, rblk=128, rblk_stride=256)
                self.add(res_b[g], bi, Cc, 1, 1, 0, ldd=1, row0=g * 2 * c_pad + 128, rblk=128, rblk_stride=256)
            if bc is not None:
                self.add(conv_b[g], bc, Cc, 1, 1, 0, ldd=1, row0=g * c_pad)
            if br is not None:
                self.add(res_b[g], br, Cc, 1, 1, 0, ldd=1, row0=g * c_pad)

    def wavenet_level_dgrad(self, convs, ress, Wp: torch.Tensor, c_pad: int):
        """pack_wavenet_level_dgrad: chain g rows g * c_pad + cin, K layout [conv_2^T | res^T | conv_0^T | conv_1^T]."""
        Cc = convs[0].shape[0]
        for g in range(len(convs)):
            self.add(convs[g], Wp, Cc, Cc, 3, 3 * Cc, taps=3, s_tap=1, row0=g * c_pad, tap_cols=c_pad, tap_pos=(2, 3, 0),
                     r_fastest=True)
            self.add(ress[g], Wp, Cc, Cc, 1, Cc, row0=g * c_pad, col0=c_pad, r_fastest=True)

    def skip_sum(self, skips, skip_b, Wp: torch.Tensor, bp: torch.Tensor, c_pad: int):
        Cc = skips[0].shape[0]
        for g, sk in enumerate(skips):
            self.add(sk, Wp, Cc, Cc, Cc, 1, col0=g * c_pad)
        srcs = list(skip_b)

        def bias_sum():
            bp[:Cc] = torch.stack([b.detach().float() for b in srcs]).sum(0)
        self.extra.append(bias_sum)

    # ------------------------------------------------------------------------------------------------ run
    def _rows(self):
        out, tile0 = [], 0
        for o in self.ops:
            tiles_r, tiles_c = -(-o["rows"] // TILE), -(-o["cols"] // TILE)
            out.append((o, tile0, tiles_c))
            tile0 += tiles_r * tiles_c
        return out, tile0

    def finalize(self, device) -> "PackTable":
        rows, total = self._rows()
        arr = (_lib.PackOp * len(rows))()
        for i, (o, tile0, tiles_c) in enumerate(rows):
            p = arr[i]
            p.src = o["src"].data_ptr() + 4 * o["src_off"]
            p.dst = o["dst"].data_ptr()
            p.s_row, p.s_col, p.s_tap = o["s_row"], o["s_col"], o["s_tap"]
            p.rows, p.cols, p.taps, p.ldd = o["rows"], o["cols"], o["taps"], o["ldd"]
            p.row0, p.rblk, p.rblk_stride = o["row0"], o["rblk"], o["rblk_stride"]
            p.col0, p.cblk, p.cblk_stride, p.tap_cols = o["col0"], o["cblk"], o["cblk_stride"], o["tap_cols"]
            for k in range(3):
                p.tap_pos[k] = o["tap_pos"][k]
            p.src_r_fastest, p.out_f32, p.tile0, p.tiles_c = int(o["r_fastest"]), int(o["out_f32"]), tile0, tiles_c
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._dev_ops = raw.to(device)
        self.total_tiles = total
        return self

    @torch.no_grad()
    def run(self):
        from .ops import _stream, check   # noqa: WPS433 (late: ops imports the library)
        for fn in self.extra:
            fn()
        check(_lib.lib.dn_pack_weights(self._dev_ops.data_ptr(), len(self.ops), self.total_tiles, _stream()), "dn_pack_weights")

    # ------------------------------------------------------------------------------------------------ reference executor
    @torch.no_grad()
    def run_reference(self):
        """The table's meaning in plain torch indexing (any device): what the tests compare dn_pack_weights against."""
        for fn in self.extra:
            fn()
        for o in self.ops:
            src = o["src"].detach().reshape(-1)[o["src_off"]:]
            r = torch.arange(o["rows"]).view(-1, 1, 1)
            c = torch.arange(o["cols"]).view(1, -1, 1)
            k = torch.arange(o["taps"]).view(1, 1, -1)
            vals = src[(r * o["s_row"] + c * o["s_col"] + k * o["s_tap"]).reshape(-1).to(src.device)]
            row = o["row0"] + (r // o["rblk"]) * o["rblk_stride"] + r % o["rblk"]
            pos = torch.tensor(o["tap_pos"][:o["taps"]]).view(1, 1, -1)
            col = o["col0"] + pos * o["tap_cols"] + (c // o["cblk"]) * o["cblk_stride"] + c % o["cblk"]
            idx = (row * o["ldd"] + col).reshape(-1).to(src.device)
            o["dst"].view(-1)[idx] = vals.to(o["dst"].dtype)
