"""DecodePipeline: the end-to-end decode with HOST tensors as an asynchronous pipeline (C ABI gd_pipeline_*, csrc/gd_pipeline.cu).

Replaces the reference's per-batch `datas = datas.to(device); pred = decoder(datas)` (quantum/decoder_v2_4.py:332-334) plus the
read-back for callers that stream batches: while batch k decodes, batch k+1 is copied in and batch k-1 is copied out.

    pipe = DecodePipeline(decoder, graph, max_B=65536, depth=3)
    t = pipe.submit_packed(prior_h, synd_h, hard_bits_out=bits_h)      # returns at once; host tensors should be pinned
    ...
    pipe.wait(t)                                                        # bits_h is complete
"""
import ctypes as C
import weakref

import torch

from . import _cabi


def _hptr(t, dtype, what):
    if t is None:
        return None
    if t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError("%s must be a contiguous %s CPU tensor" % (what, dtype))
    return C.c_void_p(t.data_ptr())


class DecodePipeline(object):
    def __init__(self, decoder, graph=None, max_B=65536, depth=3):
        g = graph or decoder._gd_graph
        if g is None:
            raise ValueError("no Tanner graph bound: call bind_graph(graph) or pass graph=")
        self.graph, self.decoder, self.max_B, self.depth = g, decoder, int(max_B), int(depth)
        self._w_host = self._pack_weights()
        handle = C.c_void_p()
        lib = _cabi.lib()
        model = decoder.gd_model()
        with torch.cuda.device(g.device):
            _cabi.check(lib.gd_pipeline_create(g.handle, C.byref(model), _hptr(self._w_host, torch.float32, "weights"), self.max_B,
                                               self.depth, C.byref(handle)), "gd_pipeline_create")
        self._h = handle
        self._finalizer = weakref.finalize(self, lib.gd_pipeline_destroy, handle)
        self._keep = {}          # ticket -> host tensors of the batch (kept alive until it is waited for / overwritten)

    def _pack_weights(self):
        params = self.decoder._gd_params()
        if not params:
            return None
        with torch.no_grad():
            return torch.cat([p.detach().reshape(-1).to("cpu", torch.float32) for p in params]).contiguous()

    def refresh_weights(self):
        """Copy the decoder's current parameters to the pipeline (in stream order: batches already submitted keep the old ones)."""
        self._w_host = self._pack_weights()
        if self._w_host is not None:
            _cabi.check(_cabi.lib().gd_pipeline_set_weights(self._h, _hptr(self._w_host, torch.float32, "weights")), "gd_pipeline_set_weights")

    def submit(self, x_host, prob_out=None, hard_out=None):
        """x_host [B, V+C] fp32 -> prob_out [B, V] fp32 and / or hard_out [B, V] uint8 (host).  Returns a ticket."""
        if prob_out is None and hard_out is None:
            raise ValueError("no output requested")
        t = C.c_int32(-1)
        _cabi.check(_cabi.lib().gd_pipeline_submit(self._h, _hptr(x_host, torch.float32, "x_host"), _hptr(prob_out, torch.float32, "prob_out"),
                                                   _hptr(hard_out, torch.uint8, "hard_out"), int(x_host.size(0)), C.byref(t)),
                    "gd_pipeline_submit")
        self._keep[t.value % (2 * self.depth)] = (x_host, prob_out, hard_out)
        return t.value

    def submit_packed(self, prior_host, synd_host, hard_bits_out=None, prob_out=None):
        """prior_host [B] fp32, synd_host [B, ceil(C/32)] int32 -> hard_bits_out [B, ceil(V/32)] int32 and / or prob_out."""
        if prob_out is None and hard_bits_out is None:
            raise ValueError("no output requested")
        t = C.c_int32(-1)
        _cabi.check(_cabi.lib().gd_pipeline_submit_packed(self._h, _hptr(prior_host, torch.float32, "prior_host"),
                                                          _hptr(synd_host, torch.int32, "synd_host"), _hptr(prob_out, torch.float32, "prob_out"),
                                                          _hptr(hard_bits_out, torch.int32, "hard_bits_out"), int(prior_host.numel()),
                                                          C.byref(t)), "gd_pipeline_submit_packed")
        self._keep[t.value % (2 * self.depth)] = (prior_host, synd_host, hard_bits_out, prob_out)
        return t.value

    def wait(self, ticket):
        _cabi.check(_cabi.lib().gd_pipeline_wait(self._h, int(ticket)), "gd_pipeline_wait")

    def drain(self):
        _cabi.check(_cabi.lib().gd_pipeline_drain(self._h), "gd_pipeline_drain")
        self._keep.clear()
