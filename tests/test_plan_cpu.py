"""CPU tests of the plan's symbolic-tensor bookkeeping (no CUDA needed)."""
from centerpose_b200.plan import Sym


def test_channel_slice_shares_storage_and_liveness():
    parent = Sym(384, 128, 128)
    parent.producer = 3
    a = Sym(64, 128, 128, parent=parent, ch_off=0)
    b = Sym(64, 128, 128, parent=parent, ch_off=320)
    assert a.pitch == 384 and b.pitch == 384 and parent.pitch == 384
    # a slice has no storage of its own and keeps the parent alive until its own last consumer
    parent.buf = object()
    assert a.buf is parent.buf and b.buf is parent.buf
    a.last_use = 10
    b.last_use = max(b.last_use, 17)
    assert parent.last_use == 17 and a.last_use == 17
    dense = Sym(64, 8, 8)
    dense.last_use = 4
    assert dense.pitch == 64 and dense.last_use == 4 and dense.buf is None
