"""Host-side runtime semantics (no GPU): dimensions, slots, operations, host arrays.

The pinned numbers are the ones the reference's own tests pin
(``test/test_accel.py:66-93,462-829``), checked here against a fake context that
records allocations instead of touching a device.
"""

import numpy as np
import pytest

from katsdpsigproc_b200 import accel
from katsdpsigproc_b200.accel import (AliasIOSlot, CompoundIOSlot, DeviceArray, Dimension,
                                      HostArray, IOSlot, Operation, OperationSequence)


class FakeBuffer:
    def __init__(self, shape, dtype, raw):
        self.shape, self.dtype, self.raw = shape, np.dtype(dtype), raw
        self.ptr = 0x1000


class FakeContext:
    """Stands in for cuda.Context: numpy memory, allocation log."""

    def __init__(self):
        self.raw_requests = []
        self.allocations = []

    def allocate_raw(self, n_bytes):
        self.raw_requests.append(n_bytes)
        return ("raw", n_bytes)

    def allocate(self, shape, dtype, raw=None):
        self.allocations.append((tuple(shape), np.dtype(dtype), raw))
        return FakeBuffer(tuple(shape), dtype, raw)

    def allocate_pinned(self, shape, dtype):
        return np.empty(shape, dtype)


class FakeQueue:
    def __init__(self, context):
        self.context = context


# ----------------------------------------------------------------------------- helpers
def test_divup_roundup():
    assert accel.divup(10, 5) == 2 and accel.divup(11, 5) == 3 and accel.divup(0, 7) == 0
    assert accel.roundup(10, 5) == 10 and accel.roundup(11, 5) == 15


# ----------------------------------------------------------------------------- HostArray
class TestHostArray:
    def setup_method(self):
        self.shape = (17, 13)
        self.padded = (20, 16)
        self.constructed = HostArray(self.shape, np.int32, self.padded)
        self.view = np.zeros(self.padded)[2:4, 3:7].view(HostArray)
        self.sliced = self.constructed[2:4, 2:4]

    def test_constructed_is_safe_and_shaped(self):
        assert HostArray.safe(self.constructed)
        assert self.constructed.shape == self.shape
        assert self.constructed.padded_shape == self.padded
        assert HostArray.padded_view(self.constructed).shape == self.padded

    def test_views_and_plain_arrays_are_not_safe(self):
        assert not HostArray.safe(self.view)
        assert not HostArray.safe(self.sliced)
        assert not HostArray.safe(np.zeros(self.shape))
        assert HostArray.padded_view(self.sliced) is None
        assert self.sliced.padded_shape == self.padded

    def test_window_is_anchored_at_the_origin(self):
        owner = HostArray.padded_view(self.constructed)
        owner[...] = 0
        self.constructed[...] = 7
        assert owner[: self.shape[0], : self.shape[1]].min() == 7
        assert owner[self.shape[0]:, :].max() == 0 and owner[:, self.shape[1]:].max() == 0

    def test_pinned_allocation_goes_through_the_context(self):
        ary = HostArray((3, 4), np.float32, (3, 8), context=FakeContext())
        assert HostArray.safe(ary) and ary.padded_shape == (3, 8)

    def test_zero_dimensional(self):
        ary = HostArray((), np.float32)
        assert HostArray.safe(ary) and ary.shape == ()


# ----------------------------------------------------------------------------- DeviceArray
class TestDeviceArrayHostSide:
    def setup_method(self):
        self.context = FakeContext()
        self.array = DeviceArray(self.context, (17, 13), np.int32, (32, 16))

    def test_properties(self):
        assert self.array.shape == (17, 13)
        assert self.array.ndim == 2
        assert self.array.dtype == np.int32
        assert self.array.padded_shape == (32, 16)
        assert self.array.strides == (64, 4)
        assert self.context.allocations == [((32, 16), np.dtype(np.int32), None)]

    def test_empty_like_and_asarray_like(self):
        like = self.array.empty_like()
        assert HostArray.safe(like) and like.shape == (17, 13) and like.padded_shape == (32, 16)
        assert self.array.asarray_like(like) is like
        plain = np.arange(17 * 13, dtype=np.int32).reshape(17, 13)
        staged = self.array.asarray_like(plain)
        assert staged is not plain and HostArray.safe(staged)
        np.testing.assert_array_equal(staged, plain)
        with pytest.raises(TypeError):
            self.array.asarray_like(plain.astype(np.float32))

    def test_region_params_merge_contiguous_axes(self):
        src = DeviceArray(self.context, (5, 6, 7), np.int16, (5, 6, 7))
        dst = DeviceArray(self.context, (5, 6, 7), np.int16, (5, 6, 7))
        so, do, shape, ss, ds = DeviceArray._region_transfer_params(src, dst, np.s_[:], np.s_[:])
        assert (so, do, shape, ss, ds) == (0, 0, (5 * 6 * 7 * 2,), (1,), (1,))

    def test_region_params_padded_and_strided(self):
        src = DeviceArray(self.context, (10, 12), np.float32, (10, 16))
        dst = DeviceArray(self.context, (4, 5), np.float32, (4, 8))
        so, do, shape, ss, ds = DeviceArray._region_transfer_params(
            src, dst, np.s_[2:6, 3:8], np.s_[:, :])
        assert so == 2 * 64 + 3 * 4 and do == 0
        assert shape == (20, 4) and ss == (1, 64) and ds == (1, 32)
        so, do, shape, ss, ds = DeviceArray._region_transfer_params(
            src, dst, np.s_[1:9:2, 7], np.s_[:, 0])
        assert so == 64 + 28 and shape == (4, 4) and ss == (1, 128) and ds == (1, 32)

    def test_region_params_newaxis_and_ints(self):
        src = DeviceArray(self.context, (3, 4), np.uint8, (3, 4))
        dst = DeviceArray(self.context, (1, 4), np.uint8, (1, 4))
        so, do, shape, ss, ds = DeviceArray._region_transfer_params(
            src, dst, np.s_[np.newaxis, 2, :], np.s_[:, :])
        assert so == 8 and shape == (4,)

    def test_region_errors(self):
        a = DeviceArray(self.context, (3, 4), np.uint8, (3, 4))
        b = DeviceArray(self.context, (3, 4), np.int8, (3, 4))
        c = DeviceArray(self.context, (3, 5), np.uint8, (3, 5))
        with pytest.raises(TypeError):
            DeviceArray._region_transfer_params(a, b, np.s_[:], np.s_[:])
        with pytest.raises(ValueError):
            DeviceArray._region_transfer_params(a, c, np.s_[:], np.s_[:])
        for bad in (np.s_[::-1], np.s_[5], np.s_[0, 0, 0], np.s_[2:2]):
            with pytest.raises(IndexError):
                DeviceArray._region_transfer_params(a, a, bad, bad)
        with pytest.raises(TypeError):
            DeviceArray._region_transfer_params(a, a, np.s_[[0, 1]], np.s_[[0, 1]])

    def test_transfer_region_peels_high_dimensions(self):
        calls = []
        DeviceArray._transfer_region(lambda *a, **k: calls.append(a), "x", "y", 0, 100,
                                     (2, 3, 4, 5), (1, 2, 6, 24), (1, 4, 12, 48))
        assert len(calls) == 5
        assert calls[2][2:] == (48, 196, (2, 3, 4), (1, 2, 6), (1, 4, 12))


# ----------------------------------------------------------------------------- Dimension
class TestDimension:
    def test_is_power2(self):
        assert all(Dimension._is_power2(v) for v in (1, 2, 32))
        assert not any(Dimension._is_power2(v) for v in (-1, 0, 3, 5))

    def test_constructor(self):
        assert Dimension(17, min_padded_round=4).min_padded_size == 20
        assert Dimension(20, min_padded_round=5).min_padded_size == 20
        with pytest.raises(ValueError):
            Dimension(10, alignment=3)
        with pytest.raises(ValueError):
            Dimension(10, min_padded_size=9)

    def test_add_align_dtype(self):
        dim = Dimension(20, alignment=8)
        assert dim.alignment == 8
        dim.add_align_dtype(np.complex64)
        assert dim.alignment_hint == 16
        dim.add_align_dtype(np.uint8)
        assert dim.alignment_hint == 128
        dim.add_align_dtype(np.float32)
        assert dim.alignment_hint == 128
        dim.add_align_dtype(np.dtype([("a", np.uint8, 3)]))   # size 3: ignored
        assert dim.alignment_hint == 128

    def test_valid(self):
        dim = Dimension(17, min_padded_round=8, alignment=4)
        assert dim.valid(24) and dim.valid(28)
        assert not dim.valid(20) and not dim.valid(30)

    def test_valid_exact(self):
        dim = Dimension(20, alignment=4, exact=True)
        assert dim.valid(20) and not dim.valid(24)
        dim = Dimension(20, min_padded_size=23, exact=True)
        assert dim.valid(23) and not dim.valid(24) and not dim.valid(20)

    @pytest.mark.parametrize("args,kwargs,expect", [
        ((30, 7), {"alignment": 4}, 36),
        ((1100, 200), {"align_dtype": np.float32}, 1216),
        ((1100,), {"align_dtype": np.float32, "exact": True}, 1100),
        ((18,), {"alignment": 8, "align_dtype": np.uint8}, 24),
        ((8320,), {"align_dtype": np.complex64}, 8320),
        ((1620,), {"align_dtype": np.uint8}, 1664),
    ])
    def test_required_padded_size(self, args, kwargs, expect):
        assert Dimension(*args, **kwargs).required_padded_size() == expect

    def test_link_merges_requirements(self):
        dim1 = Dimension(22, min_padded_size=28, alignment=4)
        dim2 = Dimension(22, min_padded_size=24, alignment=8, align_dtype=np.int32)
        dim3 = Dimension(22, min_padded_size=22, align_dtype=np.uint16)
        dim1.link(dim2)
        dim1.link(dim3)
        dim3.link(dim1)     # already linked: no-op
        for dim in (dim1, dim2, dim3):
            assert (dim.size, dim.min_padded_size, dim.alignment, dim.alignment_hint,
                    dim.exact) == (22, 28, 8, 64, False)
        dim2.add_align_dtype(np.uint8)
        assert dim3.alignment_hint == 128

    def test_link_failures_leave_both_untouched(self):
        dim1 = Dimension(22, min_padded_size=28, alignment=4)
        dim2 = Dimension(23, min_padded_size=24, alignment=8)
        with pytest.raises(ValueError):
            dim1.link(dim2)
        assert dim1._root() is not dim2._root()
        exact = Dimension(22, exact=True)
        for other in (Dimension(22, min_padded_size=28), Dimension(22, alignment=4)):
            with pytest.raises(ValueError):
                exact.link(other)
            assert exact._root() is not other._root()
            assert other.min_padded_size in (22, 28)

    def test_frozen(self):
        dim = Dimension(22)
        dim.freeze()
        assert dim.frozen
        with pytest.raises(ValueError):
            dim.add_align_dtype(np.float32)
        with pytest.raises(ValueError):
            dim.link(Dimension(22))
        with pytest.raises(ValueError):
            Dimension(22).link(dim)


# ----------------------------------------------------------------------------- IOSlot
class TestIOSlot:
    def setup_method(self):
        self.context = FakeContext()
        self.allocator = accel.DeviceAllocator(self.context)

    @pytest.mark.parametrize("bind", [True, False])
    def test_allocate(self, bind):
        dims = (Dimension(50, min_padded_round=8), Dimension(30, alignment=4))
        slot = IOSlot(dims, np.float32)
        ary = slot.allocate(self.allocator, bind=bind)
        assert ary.shape == (50, 30) and ary.padded_shape == (56, 32)
        assert (slot.buffer is ary) == bind
        assert dims[1].alignment_hint == 32          # 128 bytes of float32 on the last axis
        assert dims[0].alignment_hint == 1
        assert self.context.allocations[-1] == ((56, 32), np.dtype(np.float32), None)

    def test_one_dimensional_slots_get_no_hint(self):
        slot = IOSlot((100,), np.float32)
        assert slot.required_padded_shape() == (100,)

    def test_allocate_with_raw(self):
        slot = IOSlot((50, 30), np.uint8)
        ary = slot.allocate(self.allocator, raw="backing")
        assert slot.buffer is ary and self.context.allocations[-1][2] == "backing"

    def test_allocate_host(self):
        slot = IOSlot((Dimension(50, min_padded_round=8), Dimension(30, alignment=4)), np.float32)
        host = slot.allocate_host(self.context)
        assert HostArray.safe(host) and host.shape == (50, 30) and host.padded_shape == (56, 32)

    def test_validate(self):
        slot = IOSlot((Dimension(5, min_padded_size=8), Dimension(10, min_padded_size=10)),
                      np.float32)
        good = DeviceArray(self.context, (5, 10), np.float32, (8, 10))
        slot.validate(good)
        for shape, padded, dtype, exc in [
            ((5,), (8,), np.float32, ValueError),
            ((6, 10), (8, 10), np.float32, ValueError),
            ((5, 10), (8, 10), np.int32, TypeError),
            ((5, 10), (8, 12), np.float32, ValueError),     # more padding than required
            ((5, 10), (5, 10), np.float32, ValueError),
        ]:
            with pytest.raises(exc):
                slot.validate(DeviceArray(self.context, shape, dtype, padded))

    def test_required_bytes(self):
        slot = IOSlot((Dimension(27, alignment=4), Dimension(33, alignment=32)), np.float32)
        assert slot.required_bytes() == 4 * 28 * 64

    def test_bind_and_unbind(self):
        slot = IOSlot((Dimension(5, min_padded_size=8), 10), np.float32)
        ary = DeviceArray(self.context, (5, 10), np.float32, (8, 10))
        slot.bind(ary)
        assert slot.buffer is ary and slot.is_bound()
        assert all(d.frozen for d in slot.dimensions)
        slot.bind(None)
        assert slot.buffer is None and not slot.is_bound()

    def test_only_roots_can_be_bound(self):
        slot = IOSlot((4,), np.float32)
        CompoundIOSlot([slot])
        with pytest.raises(ValueError):
            slot.bind(None)
        with pytest.raises(ValueError):
            slot.allocate(self.allocator)


class TestCompoundIOSlot:
    def setup_method(self):
        self.context = FakeContext()
        self.dims1 = (Dimension(13, min_padded_size=17, alignment=1),
                      Dimension(7, min_padded_size=8, alignment=8),
                      Dimension(22, min_padded_size=25, alignment=4))
        self.dims2 = (Dimension(13, min_padded_size=14, alignment=4),
                      Dimension(7, min_padded_size=10, alignment=4),
                      Dimension(22, min_padded_size=22, alignment=1))
        self.slot1 = IOSlot(self.dims1, np.float32)
        self.slot2 = IOSlot(self.dims2, np.float32)

    def test_empty(self):
        with pytest.raises(ValueError):
            CompoundIOSlot([])

    def test_inconsistent_children(self):
        with pytest.raises(ValueError):
            CompoundIOSlot([self.slot1, IOSlot((13, 7, 23), np.float32)])
        with pytest.raises(TypeError):
            CompoundIOSlot([self.slot1, IOSlot((13, 7, 22), np.int32)])

    def test_combined_requirements(self):
        slot = CompoundIOSlot([self.slot1, self.slot2])
        assert slot.shape == (13, 7, 22) and slot.dtype == np.float32
        assert [d.min_padded_size for d in self.dims1] == [17, 10, 25]
        assert [d.alignment for d in self.dims1] == [4, 8, 4]
        for a, b in zip(self.dims1, self.dims2):
            assert a._root() is b._root()
        assert not self.slot1.is_root and not self.slot2.is_root and slot.is_root

    def test_bind_propagates(self):
        slot = CompoundIOSlot([self.slot1, self.slot2])
        ary = DeviceArray(self.context, (13, 7, 22), np.float32, slot.required_padded_shape())
        slot.bind(ary)
        assert slot.buffer is ary and self.slot1.buffer is ary and self.slot2.buffer is ary

    def test_bind_rejects_bad_padding(self):
        slot = CompoundIOSlot([self.slot1, self.slot2])
        with pytest.raises(ValueError):
            slot.bind(DeviceArray(self.context, (13, 7, 22), np.float32, (20, 16, 28 + 4)))
        assert slot.buffer is None and self.slot1.buffer is None and self.slot2.buffer is None

    def test_children_must_be_attachable(self):
        CompoundIOSlot([self.slot1])
        with pytest.raises(ValueError):
            CompoundIOSlot([self.slot1, self.slot2])     # slot1 already has a parent
        bound = IOSlot((13, 7, 22), np.float32)
        bound.bind(DeviceArray(self.context, (13, 7, 22), np.float32,
                               bound.required_padded_shape()))
        with pytest.raises(ValueError):
            CompoundIOSlot([bound])


class TestAliasIOSlot:
    def setup_method(self):
        self.context = FakeContext()
        self.slot1 = IOSlot((3, 7), np.float32)
        self.slot2 = IOSlot((5, 3), np.complex64)

    def test_empty(self):
        with pytest.raises(ValueError):
            AliasIOSlot([])

    def test_required_bytes(self):
        assert AliasIOSlot([self.slot1, self.slot2]).required_bytes() == 120

    def test_allocate_shares_one_raw_allocation(self):
        slot = AliasIOSlot([self.slot1, self.slot2])
        raw = slot.allocate(accel.DeviceAllocator(self.context))
        assert self.context.raw_requests == [120]
        assert slot.raw is raw and slot.is_bound()
        assert self.slot1.buffer.buffer.raw is raw and self.slot2.buffer.buffer.raw is raw

    def test_allocate_host(self):
        host = AliasIOSlot([self.slot1, self.slot2]).allocate_host(self.context)
        assert host.shape == (120,) and host.dtype == np.uint8


# ----------------------------------------------------------------------------- operations
class Leaf(Operation):
    def __init__(self, queue, log, name, slots, allocator=None):
        super().__init__(queue, allocator)
        self.log, self.name = log, name
        for slot_name, (dims, dtype) in slots.items():
            self.slots[slot_name] = IOSlot(dims, dtype)

    def _run(self):
        self.log.append(self.name)


class TestOperations:
    def setup_method(self):
        self.context = FakeContext()
        self.queue = FakeQueue(self.context)
        self.log = []

    def test_allocator_context_must_match(self):
        with pytest.raises(ValueError):
            Leaf(self.queue, self.log, "x", {}, allocator=accel.DeviceAllocator(FakeContext()))

    def test_call_binds_allocates_and_runs(self):
        op = Leaf(self.queue, self.log, "a", {"in": ((4, 6), np.float32), "out": ((4,), np.uint8)})
        given = DeviceArray(self.context, (4,), np.uint8)
        op(out=given)
        assert self.log == ["a"]
        assert op.buffer("out") is given and op.buffer("in").shape == (4, 6)
        assert op.required_bytes() == 4 * 6 * 4 + 4
        with pytest.raises(KeyError):
            op.buffer("nope")
        with pytest.raises(KeyError):
            op.bind(nope=None)

    def test_buffer_before_binding(self):
        op = Leaf(self.queue, self.log, "a", {"in": ((4,), np.float32)})
        with pytest.raises(ValueError):
            op.buffer("in")
        op.ensure_bound("in")
        assert op.buffer("in").shape == (4,)

    def test_sequence_wiring(self):
        a = Leaf(self.queue, self.log, "a", {"src": ((8, 5), np.float32), "dest": ((8, 5), np.float32)})
        b = Leaf(self.queue, self.log, "b", {"src": ((8, 5), np.float32), "tmp": ((3,), np.int16),
                                              "dest": ((5,), np.uint8)})
        seq = OperationSequence(
            self.queue, [("a", a), ("b", b)],
            compounds={"mid": ["a:dest", "b:src", "ghost:slot"], "none": ["ghost:other"]},
            aliases={"scratch": ["a:src", "b:tmp"]})
        assert set(seq.slots) == {"mid", "scratch", "b:dest"}
        assert set(seq.hidden_slots) == {"a:src", "b:tmp"}
        assert not a.is_root and not b.is_root
        seq()
        assert self.log == ["a", "b"]
        assert a.buffer("dest") is b.buffer("src") is seq.buffer("mid")
        assert seq.buffer("a:src").shape == (8, 5)       # via hidden_slots
        with pytest.raises(TypeError):
            seq.buffer("scratch")
        with pytest.raises(TypeError):
            seq.bind(scratch=None)
        assert self.context.raw_requests == [8 * 5 * 4]   # the alias slot: max of its children

    def test_sequence_rejects_foreign_children(self):
        other = Leaf(FakeQueue(self.context), self.log, "o", {})
        with pytest.raises(ValueError):
            OperationSequence(self.queue, [("o", other)])
        a = Leaf(self.queue, self.log, "a", {})
        OperationSequence(self.queue, [("a", a)])
        with pytest.raises(ValueError):
            OperationSequence(self.queue, [("a", a)])
