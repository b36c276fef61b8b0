"""Clip inference loop: the host side of `single_gpu_test` (reference tools/test_vpq.py:27-63 -- iterate the loader,
`model(return_loss=False, rescale=True, **data)`, collect the results) with the two copies a loader + collector do
around the model call made asynchronous:

  * the NEXT pair's frames are uploaded from pinned host memory on a copy stream while the current pair computes
    (what `DataLoader(pin_memory=True)` + `non_blocking` gives a user of the reference), and
  * the label maps of the finished pair are downloaded into pinned buffers on the same copy stream, so the next
    pair's kernels do not wait for the D2H.

Nothing about the model call changes: `det.simple_test` is the same entry the parity tests use."""
import torch


class ClipRunner:
    def __init__(self, det, device=None, depth=2):
        self.det = det
        self.dev = torch.device(device) if device is not None else next(det.parameters()).device
        self.copy = torch.cuda.Stream(self.dev)
        self.depth = depth
        self._out = []          # ring of pinned (pano, sem) host buffers

    def _upload(self, pair):
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy):
            a = pair[0].to(self.dev, non_blocking=True)
            b = pair[1].to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        a.record_stream(main)
        b.record_stream(main)
        return a, b, ev

    def _host_buf(self, i, like):
        while len(self._out) <= i % self.depth:
            self._out.append(None)
        buf = self._out[i % self.depth]
        if buf is None or buf[0].shape != like.shape or buf[0].dtype != like.dtype:
            buf = (torch.empty(like.shape, dtype=like.dtype).pin_memory(), torch.empty(like.shape, dtype=like.dtype).pin_memory(),
                   torch.cuda.Event())
            self._out[i % self.depth] = buf
        return buf

    def run(self, pairs, metas):
        """pairs: iterable of (img, ref_img) pinned host tensors [1,3,H,W] fp32; metas: matching img_meta dicts.
        Yields (bbox_results, segm_results, pano_results) per pair, in order; pano_results['panoptic_outputs'] and
        ['fcn_outputs'] are HOST tensors (pinned ring buffers, valid until `depth` further results were produced)."""
        main = torch.cuda.current_stream(self.dev)
        it = iter(zip(pairs, metas))
        cur = next(it, None)
        if cur is None:
            return
        nxt_up = self._upload(cur[0])
        pending = None
        i = 0
        while cur is not None:
            a, b, ev = nxt_up
            meta = cur[1]
            main.wait_event(ev)
            cur = next(it, None)
            if cur is not None:
                nxt_up = self._upload(cur[0])          # overlaps the compute below
            r = self.det.simple_test(a, [meta], ref_img=[b])
            pano, sem = r[2]["panoptic_outputs"], r[2]["fcn_outputs"]
            done = torch.cuda.Event()
            done.record(main)
            hp, hs, hev = self._host_buf(i, pano)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(done)
                hp.copy_(pano, non_blocking=True)
                hs.copy_(sem, non_blocking=True)
                hev.record(self.copy)
            pano.record_stream(self.copy)
            sem.record_stream(self.copy)
            if pending is not None:
                pending[1].synchronize()               # the previous pair's maps are on the host now
                yield pending[0]
            r[2]["panoptic_outputs"], r[2]["fcn_outputs"] = hp, hs
            pending = (r, hev)
            i += 1
        pending[1].synchronize()
        yield pending[0]
