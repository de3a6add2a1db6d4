"""Helpers for the -m gpu tests: device buffers through torch (plumbing), calls through the C ABI."""
import ctypes as C

import torch

from krep_b200 import lib
from krep_b200.abi import DeviceResult, MatchResult, Shard


def device_corpus(spec, offset, length, pad=64):
    """uint8 CUDA tensor holding corpus bytes [offset, offset+length) (offset % 16 == 0)."""
    L = lib.load()
    t = torch.empty(length + pad, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    rc = L.krep_b200_corpus_generate(C.byref(spec), t.data_ptr(), offset, length, None)
    lib.check(L)
    assert rc == 0
    return t


def to_device(data: bytes, pad=64):
    t = torch.empty(len(data) + pad, dtype=torch.uint8, device="cuda")
    if data:
        t[: len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    return t


def scan(plan, tensor, avail_len, own_begin=0, own_end=None, global_offset=0, prev_byte=-1, next_byte=-1,
         want_positions=True):
    L = lib.load()
    sh = Shard(tensor.data_ptr(), avail_len, own_begin, avail_len if own_end is None else own_end, global_offset,
               prev_byte, next_byte)
    out = DeviceResult()
    rc = L.krep_b200_scan_shard(plan, C.byref(sh), 1 if want_positions else 0, None, C.byref(out))
    lib.check(L)
    assert rc == 0
    return out


def collect(plan, params, dev):
    L = lib.load()
    res = L.krep_b200_match_result_init(16)
    try:
        cnt = L.krep_b200_collect(plan, params.ref(), C.byref(dev), res)
        lib.check(L)
        r = res.contents
        return int(cnt), [(r.positions[i].start_offset, r.positions[i].end_offset) for i in range(r.count)]
    finally:
        L.krep_b200_match_result_free(res)
