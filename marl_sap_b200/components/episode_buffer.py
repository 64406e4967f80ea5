"""EpisodeBatch / ReplayBuffer: the reference's container API on device-resident storage with kernel-side writes.

Same public surface as /root/reference/src/components/episode_buffer.py (EpisodeBatch :6-235, ReplayBuffer :237-277):
constructor arguments, ``.data.transition_data / .data.episode_data``, ``scheme / groups / batch_size /
max_seq_length / device``, ``update``, ``batch["key"]``, ``batch[("k1", "k2")]``, ``batch[bs, ts]``, ``max_t_filled``,
``to``, ``extend``; ``ReplayBuffer.insert_episode_batch / can_sample / sample`` and the ring attributes
``buffer_size / buffer_index / episodes_in_buffer``.  Error behaviour is kept too: ``KeyError`` for an unknown key in
``update``, ``ValueError`` for an unknown field name or an incompatible value shape, ``IndexError`` for a
non-contiguous time index.

What is different underneath (and why it is written differently):

* the rollout kernels write their slots straight into the ``[B, T+1, ...]`` tensors through ``kernel_view()``;
  ``update`` is off the hot path and runs as a few torch ops;
* ``ReplayBuffer`` moves whole episode rows with the ``sap_buffer_insert`` / ``sap_buffer_gather`` kernels (one launch
  per field, the ring wrap done in-kernel), and ``view_next`` lets a runner roll an episode out INSIDE the ring;
* fields named in ``lazy`` (``beta``, ``avail_actions``, ``actions_onehot``) are not stored.  ``avail_actions`` is the
  constant all-ones mask of these envs, ``actions_onehot`` is recomputed from ``actions``, and ``beta`` - the L-step
  window of the benefit tensor - is rebuilt from the benefit planes the episode was rolled out on.  For that every
  batch row carries two small per-episode integers (which plane row it read, and which GENERATION of planes), the
  planes of every generation a stored episode still refers to are kept alive by a ``BenefitSource``, and sampled or
  sliced batches rebuild their own ``beta`` with ``sap_real_beta_rows``.  So a replay buffer without a stored ``beta``
  still hands the right ``beta`` out for every episode it holds, also after the env's benefits were replaced.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch as th

from .. import _lib

PLANE_ROW, PLANE_GEN = "_plane_row", "_plane_gen"   # per-episode bookkeeping of a lazily rebuilt `beta`
_INTERNAL = (PLANE_ROW, PLANE_GEN)


def _shape_of(info, groups):
    """Per-timestep shape of a scheme entry: ``vshape`` (int or tuple), prefixed by the size of its group."""
    v = info["vshape"]
    shape = (v,) if isinstance(v, int) else tuple(v)
    group = info.get("group")
    if group:
        if group not in groups:
            raise AssertionError("Group {} must have its number of members defined in _groups_".format(group))
        shape = (groups[group],) + shape
    return shape


class BenefitSource:
    """The benefit planes behind lazily rebuilt ``beta`` fields, by generation.

    An env registers its planes whenever a rollout starts on planes it has not registered yet (``register``); episode
    rows remember the generation they were rolled out on; ``window`` rebuilds the ``beta`` rows of any batch.  Planes
    of a generation are dropped once the replay buffer that uses this source reports that no stored episode refers to
    them any more (``retain``)."""

    def __init__(self, kind, n, m, T, L):
        self.kind, self.n, self.m, self.T, self.L = kind, n, m, T, L
        self.generations = {}   # id -> (planes [rows, T, n, m], task_prios or None, shared)
        self.current = -1

    def register(self, planes, task_prios, shared):
        cur = self.generations.get(self.current)
        if cur is not None and cur[0] is planes and cur[1] is task_prios and cur[2] == bool(shared):
            return self.current
        self.current += 1
        self.generations[self.current] = (planes, task_prios, bool(shared))
        return self.current

    def holds(self, planes):
        return any(g[0] is planes for g in self.generations.values())

    def retain(self, live_generations):
        keep = set(int(g) for g in live_generations) | {self.current}
        for g in [g for g in self.generations if g not in keep]:
            del self.generations[g]

    def window(self, plane_rows, plane_gens, t0, t_count, dtype):
        """``beta[b, t]`` for the batch rows described by (plane_rows, plane_gens) and time steps [t0, t0 + t_count)."""
        B = int(plane_rows.numel())
        dev = plane_rows.device
        gens = plane_gens.tolist()   # one small host read per access of a lazy field
        inner = (self.n, self.m, self.L) if self.kind == "real" else (self.n, self.m)
        out = th.empty((B, t_count) + inner, dtype=dtype, device=dev)
        for g in sorted(set(gens)):
            if g not in self.generations:
                raise RuntimeError(f"lazy `beta`: the benefit planes of generation {g} are gone (this batch was not rolled "
                                   "out on a lazy-`beta` runner, or its source was detached); store `beta` eagerly instead")
            planes, prios, shared = self.generations[g]
            sel = th.tensor([i for i, x in enumerate(gens) if x == g], dtype=th.int64, device=dev)
            rows = plane_rows[sel].to(th.int64)
            if shared:
                rows = th.zeros_like(rows)
            if self.kind == "real":
                part = out if len(sel) == B else th.empty((len(sel), t_count) + inner, dtype=dtype, device=dev)
                dims = _lib.SapEnvDims(len(sel), self.n, self.m, self.T, self.L, 0, 0, 0)
                _lib.check(_lib.load().sap_real_beta_rows(dims, planes.data_ptr(), _lib.ptr(prios), rows.data_ptr(), t0,
                                                          t_count, part.data_ptr(), _lib.sap_dtype(dtype),
                                                          _lib.stream_ptr(dev)), "sap_real_beta_rows")
            else:   # mock env: beta[t] = S[:, :, t], zeros once the horizon is reached (mock_constellation_env.py:156-159)
                part = th.zeros((len(sel), t_count) + inner, dtype=dtype, device=dev)
                t1 = min(t0 + t_count, self.T)
                if t1 > t0:
                    part[:, :t1 - t0] = planes[rows][:, t0:t1].to(dtype)
            if part is not out:
                out[sel] = part
        return out


class EpisodeBatch:
    def __init__(self, scheme, groups, batch_size, max_seq_length, data=None, preprocess=None, device="cpu", lazy=()):
        self.scheme = scheme.copy()
        self.groups = groups
        self.batch_size = batch_size
        self.max_seq_length = max_seq_length
        self.preprocess = {} if preprocess is None else preprocess
        self.device = device
        self.lazy = frozenset(lazy)
        self.lazy_providers = {}
        self.benefit_source = None
        self._t0 = 0   # first time step of this batch inside the episode it was cut from (lazy `beta` of time slices)
        if data is None:
            self.data = self._empty_store()
            self._declare_derived_fields()
            if "filled" in scheme:
                raise AssertionError('"filled" is a reserved key for masking.')
            self.scheme["filled"] = {"vshape": (1,), "dtype": th.long}
            if "beta" in self.lazy:
                for key in _INTERNAL:
                    self.scheme[key] = {"vshape": (1,), "dtype": th.int64, "episode_const": True}
            self._allocate(self.scheme)
        else:
            self.data = data

    # ------------------------------------------------------------------ storage
    @staticmethod
    def _empty_store():
        return SimpleNamespace(transition_data={}, episode_data={})

    def _declare_derived_fields(self):
        """Scheme entries of the fields ``preprocess`` derives (``actions`` -> ``actions_onehot``): shape and dtype come
        from the transforms, group / episode_const from the source field."""
        for src, (dst, transforms) in self.preprocess.items():
            if src not in self.scheme:
                raise AssertionError(f"preprocess source {src} is not in the scheme")
            vshape, dtype = self.scheme[src]["vshape"], self.scheme[src]["dtype"]
            for tr in transforms:
                vshape, dtype = tr.infer_output_info(vshape, dtype)
            entry = {"vshape": vshape, "dtype": dtype}
            entry.update({k: self.scheme[src][k] for k in ("group", "episode_const") if k in self.scheme[src]})
            self.scheme[dst] = entry

    def _allocate(self, entries):
        for key, info in entries.items():
            if "vshape" not in info:
                raise AssertionError("Scheme must define vshape for {}".format(key))
            if key in self.lazy:
                continue
            shape = _shape_of(info, self.groups)
            dtype = info.get("dtype", th.float32)
            if info.get("episode_const", False):
                self.data.episode_data[key] = th.zeros((self.batch_size,) + shape, dtype=dtype, device=self.device)
            else:
                self.data.transition_data[key] = th.zeros((self.batch_size, self.max_seq_length) + shape, dtype=dtype,
                                                          device=self.device)

    def extend(self, scheme, groups=None):
        """Add fields after construction (episode_buffer.py:79-80)."""
        if groups is not None:
            self.groups = groups
        self.scheme.update(scheme)
        self._allocate(scheme)

    def to(self, device):
        for store in (self.data.transition_data, self.data.episode_data):
            for key in store:
                store[key] = store[key].to(device)
        self.device = device

    def field_shape(self, key):
        return _shape_of(self.scheme[key], self.groups)

    def _like(self, data, batch_size, max_seq_length, t0=0):
        """A batch over other tensors that shares this batch's scheme, lazy set and benefit source."""
        other = EpisodeBatch(self.scheme, self.groups, batch_size, max_seq_length, data=data, device=self.device, lazy=self.lazy)
        other.preprocess = self.preprocess
        other.benefit_source = self.benefit_source
        other._t0 = t0
        return other

    # ------------------------------------------------------------------ kernel-side writes
    def kernel_view(self) -> _lib.SapBatchView:
        """The C-ABI view the env kernels write through (include/marl_sap_b200.h SapBatchView)."""
        view = _lib.SapBatchView()
        stored = self.data.transition_data
        for name in _lib.VIEW_FIELDS:
            if name in stored:
                _lib.require_cuda(stored[name], name)
                setattr(view, name, _lib.field_of(stored[name]))
        staging = getattr(self, "agent_in", None)
        if staging is not None:   # [B, n, row] staging of the agent network's input: f32, or f16 rows for the split-precision fc1
            _lib.require_cuda(staging, "agent_in")
            assert staging.dtype in (th.float32, th.float16) and staging.dim() == 3 and staging.stride(2) == 1
            f = _lib.SapField()
            f.ptr, f.env_stride, f.t_stride, f.dtype = staging.data_ptr(), staging.stride(0), staging.stride(1), _lib.sap_dtype(staging.dtype)
            view.agent_in = f
        return view

    def bind_benefit_source(self, source, generation, plane_rows=None):
        """Tell a batch with a lazy ``beta`` which planes its rows are rolled out on (called by the runner per episode)."""
        if "beta" not in self.lazy:
            return
        self.benefit_source = source
        rows = self.data.episode_data[PLANE_ROW]
        if plane_rows is None:
            rows.copy_(th.arange(self.batch_size, device=rows.device).view(-1, 1))
        else:
            rows.copy_(th.as_tensor(plane_rows, device=rows.device).view(-1, 1))
        self.data.episode_data[PLANE_GEN].fill_(int(generation))

    # ------------------------------------------------------------------ update (reference :89-129)
    def update(self, data, bs=slice(None), ts=slice(None), mark_filled=True):
        index = self._index_pair((bs, ts))
        for key, value in data.items():
            if key in self.data.transition_data:
                store, where = self.data.transition_data, index
                if mark_filled:
                    store["filled"][index] = 1
                    mark_filled = False
            elif key in self.data.episode_data:
                store, where = self.data.episode_data, index[0]
            elif key in self.lazy:
                continue   # rebuilt on access: nothing to store
            else:
                raise KeyError("{} not found in transition or episode data".format(key))
            self._write(store, key, value, where)
            derived = self.preprocess.get(key)
            if derived is not None and derived[0] not in self.lazy:
                out = store[key][where]
                for tr in derived[1]:
                    out = tr.transform(out)
                dest = store[derived[0]][where]
                store[derived[0]][where] = out.to(dest.dtype).view_as(dest)

    def _write(self, store, key, value, where):
        dtype = self.scheme[key].get("dtype", th.float32)
        if isinstance(value, (list, np.ndarray)):
            value = th.tensor(np.array(value), dtype=dtype, device=self.device)   # one rounding, float64 -> field dtype
        dest = store[key][where]
        if not _trailing_compatible(value.shape, dest.shape):
            raise ValueError("Unsafe reshape of {} to {}".format(value.shape, dest.shape))
        value = value.detach() if value.device != dest.device else value
        store[key][where] = value.to(device=dest.device, dtype=dtype).view_as(dest)

    # ------------------------------------------------------------------ lazily rebuilt fields
    def set_lazy_provider(self, key, fn):
        """``fn(batch) -> full [B, T+1, ...] tensor`` for a field listed in ``lazy`` (overrides the built-in rebuild)."""
        self.lazy_providers[key] = fn

    def _rebuild(self, key):
        if key in self.lazy_providers:
            return self.lazy_providers[key](self)
        if key == "avail_actions":
            # real_constellation_env.py:267-273 / mock_constellation_env.py:205-211: every action is always available
            ones = th.ones(1, dtype=th.bool, device=self.device)
            return ones.expand((self.batch_size, self.max_seq_length) + self.field_shape(key))
        if key == "actions_onehot":
            out = self.data.transition_data["actions"]
            for tr in self.preprocess["actions"][1]:
                out = tr.transform(out)
            return out.to(self.scheme[key]["dtype"])
        if key == "beta" and self.benefit_source is not None and PLANE_ROW in self.data.episode_data:
            ep = self.data.episode_data
            return self.benefit_source.window(ep[PLANE_ROW].reshape(-1), ep[PLANE_GEN].reshape(-1), self._t0,
                                              self.max_seq_length, self.scheme["beta"]["dtype"])
        raise ValueError(f"lazy field {key} has no provider")

    # ------------------------------------------------------------------ indexing (reference :142-225)
    def __getitem__(self, item):
        if isinstance(item, str):
            return self._field(item)
        if isinstance(item, tuple) and item and all(isinstance(k, str) for k in item):
            return self._columns(item)
        return self._rows(self._index_pair(item))

    def _field(self, key):
        for store in (self.data.episode_data, self.data.transition_data):
            if key in store:
                return store[key]
        if key in self.lazy:
            return self._rebuild(key)
        print(f"key {key} not in episode or transition data")
        raise ValueError(key)

    def _columns(self, keys):
        picked = self._empty_store()
        for key in keys:
            if key in self.data.transition_data:
                picked.transition_data[key] = self.data.transition_data[key]
            elif key in self.data.episode_data:
                picked.episode_data[key] = self.data.episode_data[key]
            else:
                raise KeyError("Unrecognised key {}".format(key))
        scheme = {key: self.scheme[key] for key in keys}
        groups = {info["group"]: self.groups[info["group"]] for info in scheme.values() if "group" in info}
        return EpisodeBatch(scheme, groups, self.batch_size, self.max_seq_length, data=picked, device=self.device)

    def _rows(self, index):
        bs, ts = index
        cut = self._empty_store()
        for key, value in self.data.transition_data.items():
            cut.transition_data[key] = value[index]
        for key, value in self.data.episode_data.items():
            cut.episode_data[key] = value[bs]
        t0 = self._t0 + (ts.indices(self.max_seq_length)[0] if isinstance(ts, slice) else 0)
        out = self._like(cut, _count(bs, self.batch_size), _count(ts, self.max_seq_length), t0)
        if isinstance(ts, slice) and ts.indices(self.max_seq_length)[2] != 1:
            out.benefit_source = None   # a strided time slice has no contiguous window to rebuild
        parent = self
        for key, fn in self.lazy_providers.items():
            out.lazy_providers[key] = (lambda _b, fn=fn: fn(parent)[index])
        return out

    @staticmethod
    def _index_pair(item):
        """(batch index, time index); ints become length-1 slices, a bare batch index selects every time step."""
        if isinstance(item, (slice, int, list, np.ndarray, th.Tensor)):
            item = (item, slice(None))
        if isinstance(item[1], list):
            raise IndexError("Indexing across Time must be contiguous")
        return tuple(slice(i, i + 1) if isinstance(i, int) else i for i in item)

    def max_t_filled(self):
        return th.sum(self.data.transition_data["filled"], 1).max(0)[0]

    def __repr__(self):
        return "EpisodeBatch. Batch Size:{} Max_seq_len:{} Keys:{} Groups:{}".format(
            self.batch_size, self.max_seq_length, self.scheme.keys(), self.groups.keys())


def _trailing_compatible(src, dst):
    """``view_as`` is safe when, matching dims from the right, every dim of ``dst`` is either the next dim of ``src`` or 1."""
    i = len(src) - 1
    for d in reversed(dst):
        if i >= 0 and src[i] == d:
            i -= 1
        elif d != 1:
            return False
    return True


def _count(index, size):
    if isinstance(index, slice):
        return len(range(*index.indices(size)))
    if isinstance(index, th.Tensor):
        return int(index.numel())
    return len(index)


class ReplayBuffer(EpisodeBatch):
    def __init__(self, scheme, groups, buffer_size, max_seq_length, preprocess=None, device="cpu", lazy=()):
        super().__init__(scheme, groups, buffer_size, max_seq_length, preprocess=preprocess, device=device, lazy=lazy)
        self.buffer_size = buffer_size
        self.buffer_index = 0
        self.episodes_in_buffer = 0
        self.kernel_launches = 0   # sap_buffer_insert / sap_buffer_gather launches issued so far

    # ------------------------------------------------------------------ ring insert (reference :244-259)
    def view_next(self, n_new):
        """An EpisodeBatch whose tensors ALIAS the next ``n_new`` ring rows, for runners that roll out straight into the
        replay buffer (no episode-sized copy at insert time).  ``None`` when the rows would wrap.  Rolling into a used
        slot is safe because the env kernels rewrite every slot they own each episode and the slots they never write
        (actions / rewards / terminated at t = T) are zero from allocation."""
        if n_new > self.buffer_size or self.buffer_index + n_new > self.buffer_size:
            return None
        view = self._rows((slice(self.buffer_index, self.buffer_index + n_new), slice(None)))
        view._ring_owner, view._ring_start = self, self.buffer_index
        return view

    def _advance(self, n_new):
        end = self.buffer_index + n_new
        self.episodes_in_buffer = max(self.episodes_in_buffer, min(end, self.buffer_size))
        self.buffer_index = end % self.buffer_size

    def insert_episode_batch(self, ep_batch):
        n_new = ep_batch.batch_size
        if n_new > self.buffer_size:
            raise ValueError(f"episode batch of {n_new} does not fit a replay buffer of {self.buffer_size}")
        if ep_batch.benefit_source is not None:
            self.benefit_source = ep_batch.benefit_source
        owner = getattr(ep_batch, "_ring_owner", None)
        if owner is self:
            if ep_batch._ring_start != self.buffer_index:
                # a view of OTHER ring rows (the ring moved on since view_next): its storage aliases ours, and the row copy
                # below must not read and write overlapping ranges -> go through a private copy
                ep_batch = self._detached_copy(ep_batch)
            else:   # rolled out in place: only the ring bookkeeping moves
                self._advance(n_new)
                self._prune_generations()
                return
        if self._rows_copyable(ep_batch):
            lib = _lib.load()
            stream = _lib.stream_ptr(self.data.transition_data["filled"].device)
            for name in ("transition_data", "episode_data"):
                mine, theirs = getattr(self.data, name), getattr(ep_batch.data, name)
                for key, src in theirs.items():
                    if key not in mine:
                        continue   # a field this buffer rebuilds lazily (update() skips it the same way)
                    dst = mine[key]
                    row_bytes = dst[0].numel() * dst.element_size()
                    _lib.check(lib.sap_buffer_insert(dst.data_ptr(), src.data_ptr(), row_bytes, self.buffer_size,
                                                     self.buffer_index, 0, n_new, stream), "sap_buffer_insert")
                    self.kernel_launches += 1
            self._advance(n_new)
        else:   # host tensors or views: the same ring semantics with torch copies, split where the ring wraps
            first = min(n_new, self.buffer_size - self.buffer_index)
            for lo, hi in ((0, first), (first, n_new)):
                if hi > lo:
                    part = ep_batch[lo:hi, :]
                    rows = slice(self.buffer_index, self.buffer_index + (hi - lo))
                    self.update(part.data.transition_data, rows, slice(0, part.max_seq_length), mark_filled=False)
                    self.update(part.data.episode_data, rows)
                    self._advance(hi - lo)
        self._prune_generations()

    def _detached_copy(self, batch):
        copy = self._empty_store()
        for name in ("transition_data", "episode_data"):
            for key, value in getattr(batch.data, name).items():
                getattr(copy, name)[key] = value.clone()
        return batch._like(copy, batch.batch_size, batch.max_seq_length, batch._t0)

    def _rows_copyable(self, ep_batch):
        """The whole-row kernel copy applies when both sides are CUDA, contiguous and shaped alike."""
        if ep_batch.max_seq_length != self.max_seq_length:
            return False
        for name in ("transition_data", "episode_data"):
            mine, theirs = getattr(self.data, name), getattr(ep_batch.data, name)
            for key, src in theirs.items():
                if key not in mine:
                    if key in self.lazy or key in _INTERNAL:
                        continue
                    raise KeyError("{} not found in transition or episode data".format(key))
                dst = mine[key]
                if not (src.is_cuda and dst.is_cuda and src.device == dst.device and src.is_contiguous() and dst.is_contiguous()
                        and src.dtype == dst.dtype and src.shape[1:] == dst.shape[1:]):
                    return False
        return True

    def _prune_generations(self):
        if self.benefit_source is not None and PLANE_GEN in self.data.episode_data and len(self.benefit_source.generations) > 1:
            live = self.data.episode_data[PLANE_GEN][:self.episodes_in_buffer].unique().tolist()
            self.benefit_source.retain(live)

    def can_sample(self, batch_size):
        return self.episodes_in_buffer >= batch_size

    # ------------------------------------------------------------------ sample (reference :264-271)
    def sample(self, batch_size):
        assert self.can_sample(batch_size)
        if self.episodes_in_buffer == batch_size:
            return self[:batch_size]
        # uniform without replacement, host RNG like the reference
        return self.gather(np.random.choice(self.episodes_in_buffer, batch_size, replace=False))

    def gather(self, ep_ids):
        """A copy of the chosen episodes (the reference's fancy-index ``self[ep_ids]``)."""
        probe = self.data.transition_data["filled"]
        if not probe.is_cuda:
            return self[ep_ids]
        lib = _lib.load()
        ids = th.as_tensor(np.asarray(ep_ids), dtype=th.int64, device=probe.device)
        count = int(ids.numel())
        picked = self._empty_store()
        stream = _lib.stream_ptr(probe.device)
        for name in ("transition_data", "episode_data"):
            for key, src in getattr(self.data, name).items():
                dst = th.empty((count,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
                row_bytes = src[0].numel() * src.element_size()
                _lib.check(lib.sap_buffer_gather(dst.data_ptr(), src.data_ptr(), ids.data_ptr(), row_bytes, count, stream),
                           "sap_buffer_gather")
                self.kernel_launches += 1
                getattr(picked, name)[key] = dst
        return self._like(picked, count, self.max_seq_length)

    def __repr__(self):
        return "ReplayBuffer. {}/{} episodes. Keys:{} Groups:{}".format(
            self.episodes_in_buffer, self.buffer_size, self.scheme.keys(), self.groups.keys())
