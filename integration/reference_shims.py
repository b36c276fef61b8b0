"""Reference-side bindings a maintainer of mcahny/vps would add to swap single native ops for libvps_b200.so while keeping
the reference's Python modules.  Raw ctypes against the C ABI of include/vps_b200.h -- no
vps_b200 Python import, exactly what would live next to the reference's own wrappers.  Executed by
tests/test_gpu_shims.py against the oracle of the op each one replaces.

  nms_cuda_nms(dets, thr)              replaces  nms_cuda.nms                 (mmdet/ops/nms/nms_wrapper.py:43, src/nms_cuda.cpp:8-16)
  resample2d_forward(in1, flow, out)   replaces  resample2d_cuda.forward     (resample2d_package/resample2d.py:18-19)
  channelnorm_forward(in1, out)        replaces  channelnorm_cuda.forward    (channelnorm_package/channelnorm.py:15)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.environ.get("VPS_B200_LIB", os.path.join(_HERE, "..", "vps_b200", "lib", "libvps_b200.so")))
_lib.vps_last_error.restype = ctypes.c_char_p


class _T(ctypes.Structure):       # vps_tensor (include/vps_b200.h)
    _fields_ = [("ptr", ctypes.c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("cs", ctypes.c_int32), ("dtype", ctypes.c_int32)]


def _ok(status, what):
    if status != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, status, _lib.vps_last_error().decode()))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _nhwc(t_nchw):
    """the reference's ops are NCHW fp32; the library is NHWC: one permuted copy each way at this boundary"""
    x = t_nchw.permute(0, 2, 3, 1).contiguous()
    n, h, w, c = x.shape
    return x, _T(x.data_ptr(), n, h, w, c, c, 0)


def nms_cuda_nms(dets, thr):
    """same contract as nms_cuda.nms: dets [N,5] fp32 CUDA (x1,y1,x2,y2,score) -> LongTensor of kept indices, ascending"""
    n = dets.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.long, device=dets.device)
    s = _stream()
    dets = dets.contiguous().float()
    scores = dets[:, 4].contiguous()
    ks, idx = torch.empty_like(scores), torch.empty(n, dtype=torch.int32, device=dets.device)
    ws = torch.empty(n * 24 + 65536, dtype=torch.uint8, device=dets.device)
    _ok(_lib.vps_sort_desc(ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ks.data_ptr()), ctypes.c_void_p(idx.data_ptr()), n,
                           ctypes.c_void_p(ws.data_ptr()), ctypes.c_int64(ws.numel()), s), "vps_sort_desc")
    sorted_dets = torch.empty_like(dets)
    _ok(_lib.vps_gather_rows(ctypes.c_void_p(dets.data_ptr()), ctypes.c_void_p(idx.data_ptr()), n, None, 5,
                             ctypes.c_void_p(sorted_dets.data_ptr()), s), "vps_gather_rows")
    keep = torch.empty(n, dtype=torch.int32, device=dets.device)
    nk = torch.zeros(1, dtype=torch.int32, device=dets.device)
    mws = torch.empty(max(n * ((n + 63) // 64) * 8, 8), dtype=torch.uint8, device=dets.device)
    _ok(_lib.vps_nms(ctypes.c_void_p(sorted_dets.data_ptr()), n, None, ctypes.c_float(thr), ctypes.c_void_p(keep.data_ptr()),
                     ctypes.c_void_p(nk.data_ptr()), ctypes.c_void_p(mws.data_ptr()), ctypes.c_int64(mws.numel()), s), "vps_nms")
    return idx[keep[: int(nk.item())].long()].long().sort()[0]


def resample2d_forward(in1, flow, out):
    """resample2d_cuda.forward(input1, input2, output, kernel_size=1, bilinear=True): NCHW fp32 CUDA tensors"""
    x, tx = _nhwc(in1.float())
    f, tf = _nhwc(flow.float())
    y = torch.empty(out.shape[0], out.shape[2], out.shape[3], out.shape[1], dtype=torch.float32, device=out.device)
    ty = _T(y.data_ptr(), y.shape[0], y.shape[1], y.shape[2], y.shape[3], y.shape[3], 0)
    _ok(_lib.vps_resample2d(ctypes.byref(tx), ctypes.byref(tf), ctypes.byref(ty), _stream()), "vps_resample2d")
    out.copy_(y.permute(0, 3, 1, 2))
    return out


def channelnorm_forward(in1, out):
    """channelnorm_cuda.forward(input1, output, norm_deg=2): out[b,0,y,x] = sqrt(sum_c in1^2)"""
    x, tx = _nhwc(in1.float())
    y = torch.empty(out.shape[0], out.shape[2], out.shape[3], 1, dtype=torch.float32, device=out.device)
    ty = _T(y.data_ptr(), y.shape[0], y.shape[1], y.shape[2], 1, 1, 0)
    _ok(_lib.vps_channelnorm(ctypes.byref(tx), None, ctypes.byref(ty), _stream()), "vps_channelnorm")
    out.copy_(y.permute(0, 3, 1, 2))
    return out
