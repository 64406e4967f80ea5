"""EpisodeBatch / ReplayBuffer with the reference's API, device-resident storage and kernel-side writes.

Mirrors /root/reference/src/components/episode_buffer.py:
  EpisodeBatch  :6-235  (_setup_data :30-77, update :89-129, __getitem__ :142-182, max_t_filled :227-228)
  ReplayBuffer  :237-277 (insert_episode_batch :244-259, can_sample :261-262, sample :264-271)

Differences that matter on a B200:
  * the rollout kernels write their slots directly into the [B, T+1, ...] tensors (``kernel_view``),
    so ``update`` is not on the hot path; it is kept for API compatibility and runs as torch ops;
  * ``ReplayBuffer.insert_episode_batch`` / ``sample`` move whole episode rows with the
    ``sap_buffer_insert`` / ``sap_buffer_gather`` kernels (one launch per field, ring wrap in-kernel);
  * fields listed in ``lazy`` (``beta``, ``avail_actions``, ``actions_onehot``) are not stored; they are
    rebuilt on access from the benefit planes / the actions.
"""
from __future__ import annotations

from types import SimpleNamespace as SN

import numpy as np
import torch as th

from .. import _lib


def _as_tuple(vshape):
    return (vshape,) if isinstance(vshape, int) else tuple(vshape)


class EpisodeBatch:
    def __init__(self, scheme, groups, batch_size, max_seq_length, data=None, preprocess=None, device="cpu",
                 lazy=()):
        self.scheme = scheme.copy()
        self.groups = groups
        self.batch_size = batch_size
        self.max_seq_length = max_seq_length
        self.preprocess = {} if preprocess is None else preprocess
        self.device = device
        self.lazy = frozenset(lazy)
        self.lazy_providers = {}
        if data is not None:
            self.data = data
        else:
            self.data = SN()
            self.data.transition_data = {}
            self.data.episode_data = {}
            self._setup_data(self.scheme, self.groups, batch_size, max_seq_length, self.preprocess)

    # ------------------------------------------------------------------ allocation (:30-77)
    def _setup_data(self, scheme, groups, batch_size, max_seq_length, preprocess):
        if preprocess is not None:
            for k, (new_k, transforms) in preprocess.items():
                assert k in scheme
                vshape, dtype = self.scheme[k]["vshape"], self.scheme[k]["dtype"]
                for tr in transforms:
                    vshape, dtype = tr.infer_output_info(vshape, dtype)
                self.scheme[new_k] = {"vshape": vshape, "dtype": dtype}
                for inherit in ("group", "episode_const"):
                    if inherit in self.scheme[k]:
                        self.scheme[new_k][inherit] = self.scheme[k][inherit]
        assert "filled" not in scheme, '"filled" is a reserved key for masking.'
        scheme.update({"filled": {"vshape": (1,), "dtype": th.long}})

        for key, info in scheme.items():
            assert "vshape" in info, "Scheme must define vshape for {}".format(key)
            shape = _as_tuple(info["vshape"])
            group = info.get("group", None)
            if group:
                assert group in groups, "Group {} must have its number of members defined in _groups_".format(group)
                shape = (groups[group], *shape)
            if key in self.lazy:
                continue
            dtype = info.get("dtype", th.float32)
            if info.get("episode_const", False):
                self.data.episode_data[key] = th.zeros((batch_size, *shape), dtype=dtype, device=self.device)
            else:
                self.data.transition_data[key] = th.zeros((batch_size, max_seq_length, *shape), dtype=dtype,
                                                          device=self.device)

    def extend(self, scheme, groups=None):
        self._setup_data(scheme, self.groups if groups is None else groups, self.batch_size, self.max_seq_length, None)

    def to(self, device):
        for store in (self.data.transition_data, self.data.episode_data):
            for k, v in store.items():
                store[k] = v.to(device)
        self.device = device

    def field_shape(self, key):
        info = self.scheme[key]
        shape = _as_tuple(info["vshape"])
        if info.get("group"):
            shape = (self.groups[info["group"]], *shape)
        return shape

    # ------------------------------------------------------------------ kernel-side writes
    def kernel_view(self) -> _lib.SapBatchView:
        """The C-ABI view the fused step kernels write through (include/marl_sap_b200.h SapBatchView)."""
        view = _lib.SapBatchView()
        td = self.data.transition_data
        for name in _lib.VIEW_FIELDS:
            if name in td:
                _lib.require_cuda(td[name], name)
                setattr(view, name, _lib.field_of(td[name]))
        ai = getattr(self, "agent_in", None)
        if ai is not None:  # [B, n, row] staging of the agent network's input (see SapBatchView.agent_in): f32, or f16
            _lib.require_cuda(ai, "agent_in")   # rows (packed or padded to a multiple of 8 columns) for the split-precision fc1
            assert ai.dtype in (th.float32, th.float16) and ai.dim() == 3 and ai.stride(2) == 1
            f = _lib.SapField()
            f.ptr, f.env_stride, f.t_stride, f.dtype = ai.data_ptr(), ai.stride(0), ai.stride(1), _lib.sap_dtype(ai.dtype)
            view.agent_in = f
        return view

    # ------------------------------------------------------------------ update (:89-129)
    def update(self, data, bs=slice(None), ts=slice(None), mark_filled=True):
        slices = self._parse_slices((bs, ts))
        for k, v in data.items():
            if k in self.data.transition_data:
                target = self.data.transition_data
                if mark_filled:
                    target["filled"][slices] = 1
                    mark_filled = False
                _slices = slices
            elif k in self.data.episode_data:
                target = self.data.episode_data
                _slices = slices[0]
            elif k in self.lazy:
                continue  # rebuilt on access, nothing to store
            else:
                raise KeyError("{} not found in transition or episode data".format(k))

            dtype = self.scheme[k].get("dtype", th.float32)
            if type(v) == list:
                v = th.tensor(np.array(v), dtype=dtype, device=self.device)  # single rounding fp64 -> dtype
            dest = target[k][_slices]
            self._check_safe_view(v, dest)
            if v.device != dest.device:
                v = v.detach().to(dest.device)
            if v.dtype != dtype:
                v = v.to(dtype)
            target[k][_slices] = v.view_as(dest)

            if k in self.preprocess:
                new_k = self.preprocess[k][0]
                if new_k in self.lazy:
                    continue
                v = target[k][_slices]
                for transform in self.preprocess[k][1]:
                    v = transform.transform(v)
                v = v.to(dtype)
                target[new_k][_slices] = v.view_as(target[new_k][_slices])

    @staticmethod
    def _check_safe_view(v, dest):
        idx = len(v.shape) - 1
        for s in dest.shape[::-1]:
            if v.shape[idx] != s:
                if s != 1:
                    raise ValueError("Unsafe reshape of {} to {}".format(v.shape, dest.shape))
            else:
                idx -= 1

    # ------------------------------------------------------------------ lazily rebuilt fields
    def set_lazy_provider(self, key, fn):
        """fn(batch) -> full [B, T+1, ...] tensor for a field listed in ``lazy``."""
        self.lazy_providers[key] = fn

    def _lazy_field(self, key):
        if key == "avail_actions" and key not in self.lazy_providers:
            # real_constellation_env.py:267-273 / mock_constellation_env.py:205-211: every action is always available
            shape = self.field_shape(key)
            return th.ones(1, dtype=th.bool, device=self.device).expand(self.batch_size, self.max_seq_length, *shape)
        if key == "actions_onehot" and key not in self.lazy_providers:
            acts = self.data.transition_data["actions"]
            v = acts
            for transform in self.preprocess["actions"][1]:
                v = transform.transform(v)
            return v.to(self.scheme[key]["dtype"])
        if key in self.lazy_providers:
            return self.lazy_providers[key](self)
        raise ValueError(f"lazy field {key} has no provider")

    # ------------------------------------------------------------------ indexing (:142-225)
    def __getitem__(self, item):
        if isinstance(item, str):
            if item in self.data.episode_data:
                return self.data.episode_data[item]
            elif item in self.data.transition_data:
                return self.data.transition_data[item]
            elif item in self.lazy:
                return self._lazy_field(item)
            else:
                print(f"key {item} not in episode or transition data")
                raise ValueError
        elif isinstance(item, tuple) and all([isinstance(it, str) for it in item]):
            new_data = self._new_data_sn()
            for key in item:
                if key in self.data.transition_data:
                    new_data.transition_data[key] = self.data.transition_data[key]
                elif key in self.data.episode_data:
                    new_data.episode_data[key] = self.data.episode_data[key]
                else:
                    raise KeyError("Unrecognised key {}".format(key))
            new_scheme = {key: self.scheme[key] for key in item}
            new_groups = {self.scheme[key]["group"]: self.groups[self.scheme[key]["group"]]
                          for key in item if "group" in self.scheme[key]}
            return EpisodeBatch(new_scheme, new_groups, self.batch_size, self.max_seq_length, data=new_data,
                                device=self.device)
        else:
            item = self._parse_slices(item)
            new_data = self._new_data_sn()
            for k, v in self.data.transition_data.items():
                new_data.transition_data[k] = v[item]
            for k, v in self.data.episode_data.items():
                new_data.episode_data[k] = v[item[0]]
            ret_bs = self._get_num_items(item[0], self.batch_size)
            ret_max_t = self._get_num_items(item[1], self.max_seq_length)
            ret = EpisodeBatch(self.scheme, self.groups, ret_bs, ret_max_t, data=new_data, device=self.device,
                               lazy=self.lazy)
            ret.preprocess = self.preprocess
            if self.lazy_providers:
                parent, sl = self, item
                for key, fn in self.lazy_providers.items():
                    ret.lazy_providers[key] = (lambda _b, fn=fn: fn(parent)[sl])
            return ret

    @staticmethod
    def _get_num_items(indexing_item, max_size):
        if isinstance(indexing_item, (list, np.ndarray)):
            return len(indexing_item)
        if isinstance(indexing_item, th.Tensor):
            return int(indexing_item.numel())
        if isinstance(indexing_item, slice):
            _range = indexing_item.indices(max_size)
            return 1 + (_range[1] - _range[0] - 1) // _range[2]

    @staticmethod
    def _new_data_sn():
        new_data = SN()
        new_data.transition_data = {}
        new_data.episode_data = {}
        return new_data

    @staticmethod
    def _parse_slices(items):
        parsed = []
        if isinstance(items, (slice, int, list, np.ndarray, th.Tensor)):
            items = (items, slice(None))
        if isinstance(items[1], list):
            raise IndexError("Indexing across Time must be contiguous")
        for item in items:
            if isinstance(item, int):
                parsed.append(slice(item, item + 1))
            else:
                parsed.append(item)
        return tuple(parsed)

    def max_t_filled(self):
        return th.sum(self.data.transition_data["filled"], 1).max(0)[0]

    def __repr__(self):
        return "EpisodeBatch. Batch Size:{} Max_seq_len:{} Keys:{} Groups:{}".format(
            self.batch_size, self.max_seq_length, self.scheme.keys(), self.groups.keys())


class ReplayBuffer(EpisodeBatch):
    def __init__(self, scheme, groups, buffer_size, max_seq_length, preprocess=None, device="cpu", lazy=()):
        super().__init__(scheme, groups, buffer_size, max_seq_length, preprocess=preprocess, device=device, lazy=lazy)
        self.buffer_size = buffer_size
        self.buffer_index = 0
        self.episodes_in_buffer = 0
        self.kernel_launches = 0  # sap_buffer_insert / sap_buffer_gather launches issued so far

    # ------------------------------------------------------------------ ring insert (:244-259)
    def view_next(self, n_new):
        """An EpisodeBatch whose tensors ALIAS the next ``n_new`` ring rows, for runners that roll out straight into
        the replay buffer (no episode-sized copy at insert time).  ``None`` when the rows would wrap.  Rolling into
        a used slot is safe because the env kernels rewrite every slot they own each episode and the slots they
        never write (actions/rewards/terminated at t = T) are zero from allocation."""
        if n_new > self.buffer_size or self.buffer_index + n_new > self.buffer_size:
            return None
        sl = slice(self.buffer_index, self.buffer_index + n_new)
        data = self._new_data_sn()
        for k, v in self.data.transition_data.items():
            data.transition_data[k] = v[sl]
        for k, v in self.data.episode_data.items():
            data.episode_data[k] = v[sl]
        view = EpisodeBatch(self.scheme, self.groups, n_new, self.max_seq_length, data=data, device=self.device,
                            lazy=self.lazy)
        view.preprocess = self.preprocess
        view._ring_owner, view._ring_start = self, self.buffer_index
        return view

    def insert_episode_batch(self, ep_batch):
        n_new = ep_batch.batch_size
        if n_new > self.buffer_size:
            raise ValueError(f"episode batch of {n_new} does not fit a replay buffer of {self.buffer_size}")
        if getattr(ep_batch, "_ring_owner", None) is self and ep_batch._ring_start == self.buffer_index:
            # the episode was rolled out in place (view_next): only the ring bookkeeping moves
            end = self.buffer_index + n_new
            self.episodes_in_buffer = max(self.episodes_in_buffer, end)
            self.buffer_index = end % self.buffer_size
            return
        if self._device_rows_ok(ep_batch):
            lib = _lib.load()
            stream = _lib.stream_ptr(self.data.transition_data["filled"].device)
            for store_name in ("transition_data", "episode_data"):
                dst_store, src_store = getattr(self.data, store_name), getattr(ep_batch.data, store_name)
                for k, src in src_store.items():
                    dst = dst_store[k]
                    row_bytes = dst[0].numel() * dst.element_size()
                    _lib.check(lib.sap_buffer_insert(dst.data_ptr(), src.data_ptr(), row_bytes, self.buffer_size,
                                                     self.buffer_index, 0, n_new, stream), "sap_buffer_insert")
                    self.kernel_launches += 1
            end = self.buffer_index + n_new
            self.episodes_in_buffer = max(self.episodes_in_buffer, min(end, self.buffer_size))
            self.buffer_index = end % self.buffer_size
            return
        # host-side containers (CPU tensors / views): same ring semantics with torch copies
        if self.buffer_index + n_new <= self.buffer_size:
            self.update(ep_batch.data.transition_data, slice(self.buffer_index, self.buffer_index + n_new),
                        slice(0, ep_batch.max_seq_length), mark_filled=False)
            self.update(ep_batch.data.episode_data, slice(self.buffer_index, self.buffer_index + n_new))
            self.buffer_index = self.buffer_index + n_new
            self.episodes_in_buffer = max(self.episodes_in_buffer, self.buffer_index)
            self.buffer_index = self.buffer_index % self.buffer_size
            assert self.buffer_index < self.buffer_size
        else:
            buffer_left = self.buffer_size - self.buffer_index
            self.insert_episode_batch(ep_batch[0:buffer_left, :])
            self.insert_episode_batch(ep_batch[buffer_left:, :])

    def _device_rows_ok(self, ep_batch):
        """Whole-row kernel copy is valid when both sides are CUDA, contiguous and shaped alike."""
        if ep_batch.max_seq_length != self.max_seq_length:
            return False
        for store_name in ("transition_data", "episode_data"):
            dst_store, src_store = getattr(self.data, store_name), getattr(ep_batch.data, store_name)
            for k, src in src_store.items():
                if k not in dst_store:
                    raise KeyError("{} not found in transition or episode data".format(k))
                dst = dst_store[k]
                if not (src.is_cuda and dst.is_cuda and src.device == dst.device and src.is_contiguous()
                        and dst.is_contiguous() and src.dtype == dst.dtype and src.shape[1:] == dst.shape[1:]):
                    return False
        return True

    def can_sample(self, batch_size):
        return self.episodes_in_buffer >= batch_size

    # ------------------------------------------------------------------ sample (:264-271)
    def sample(self, batch_size):
        assert self.can_sample(batch_size)
        if self.episodes_in_buffer == batch_size:
            return self[:batch_size]
        ep_ids = np.random.choice(self.episodes_in_buffer, batch_size, replace=False)  # uniform, host RNG like the reference
        return self.gather(ep_ids)

    def gather(self, ep_ids):
        """A copy of the chosen episodes (the reference's fancy-index ``self[ep_ids]``)."""
        some = self.data.transition_data["filled"]
        if not some.is_cuda:
            return self[ep_ids]
        lib = _lib.load()
        ids = th.as_tensor(np.asarray(ep_ids), dtype=th.int64, device=some.device)
        count = int(ids.numel())
        new_data = self._new_data_sn()
        stream = _lib.stream_ptr(some.device)
        for store_name in ("transition_data", "episode_data"):
            for k, src in getattr(self.data, store_name).items():
                dst = th.empty((count, *src.shape[1:]), dtype=src.dtype, device=src.device)
                row_bytes = src[0].numel() * src.element_size()
                _lib.check(lib.sap_buffer_gather(dst.data_ptr(), src.data_ptr(), ids.data_ptr(), row_bytes, count, stream),
                           "sap_buffer_gather")
                self.kernel_launches += 1
                getattr(new_data, store_name)[k] = dst
        ret = EpisodeBatch(self.scheme, self.groups, count, self.max_seq_length, data=new_data, device=self.device,
                           lazy=self.lazy)
        ret.preprocess = self.preprocess
        return ret

    def __repr__(self):
        return "ReplayBuffer. {}/{} episodes. Keys:{} Groups:{}".format(
            self.episodes_in_buffer, self.buffer_size, self.scheme.keys(), self.groups.keys())
