"""Input pipeline of the UniVTG hot path: packed 16-bit feature shards + a pinned, double-buffered batch loader
(SURVEY.md section 8 row f-2).

What it replaces in the reference (per SAMPLE and per step: two or three np.load of small .npz files, L2 normalisation, TEF
concatenation, zero padding to the batch maximum, fp32 host tensors, a blocking H2D copy):
  main/dataset.py:644-696   _get_query_feat_by_qid / _get_video_feat_by_vid  ({qid}.npz['last_hidden_state'], {vid}.npz['features'],
                            utils/basic_utils.py:97-99 l2_normalize_np_array, truncation to the shortest feature type)
  main/dataset.py:534-540   TEF columns [l / L_v, (l + 1) / L_v]
  main/dataset.py:1037-1052 start_end_collate_mr -> utils/tensor_utils.py:6-53 pad_sequences_1d (pad to the batch max, float32 masks)
  main/dataset.py:1071-1077 prepare_batch_inputs_mr (H2D)
  data/create_h5py.py:19-35 (the reference's own attempt at a packed cache)

Here the per-sample work happens ONCE, at packing time: a shard file holds every video / query feature matrix already normalised,
TEF-extended and rounded to fp16 (2 bytes per element on disk, in pinned memory, over PCIe and in the first LayerNorm's read), plus
offset tables.  `ShardLoader` memory-maps the shard, assembles each padded batch into pinned staging buffers with a small thread
pool, and copies it to the device on a side stream one batch ahead of the consumer.  The model consumes the fp16 tensors directly
(univtg_plan_set_input_format); no arithmetic happens here - only copies.

File layout (little endian):  b"UVSHARD1" | u64 header_bytes | header JSON (padded) | arrays at 64-byte aligned offsets named in the
header: vid_off i64 [n_vid + 1] (row offsets), txt_off i64 [n_txt + 1], samples i32 [n, 2] (video index, query index),
vid f16 [rows_v, Dv], txt f16 [rows_t, Dt].
"""
import json
import os
import threading

import numpy as np
import torch

MAGIC = b"UVSHARD1"


def l2_normalize(x, eps=1e-5):
    """utils/basic_utils.py:97-99."""
    return x / (np.linalg.norm(x, axis=-1, keepdims=True) + eps)


def prepare_video(feature_list, normalize=True, use_tef=True, max_v_l=None):
    """One video's clip features as the reference feeds them (main/dataset.py:674-690, 534-540): every feature type L2-normalised
    per clip, cut to the shortest type (and to max_v_l), concatenated, TEF columns appended.  -> float32 [L_v, sum(D_i) + 2]."""
    feats = [np.asarray(f, dtype=np.float32) for f in feature_list]
    if normalize:
        feats = [l2_normalize(f) for f in feats]
    n = min(len(f) for f in feats)
    if max_v_l is not None:
        n = min(n, int(max_v_l))
    v = np.concatenate([f[:n] for f in feats], axis=1)
    if use_tef:
        st = np.arange(0, n, 1.0, dtype=np.float32) / n
        v = np.concatenate([v, np.stack([st, st + np.float32(1.0) / n], axis=1).astype(np.float32)], axis=1)
    return v


def prepare_query(feat, normalize=True):
    """main/dataset.py:652-664: token features, L2-normalised per token (no truncation: the reference commented it out)."""
    q = np.asarray(feat, dtype=np.float32)
    return l2_normalize(q) if normalize else q


def write_shard(path, videos, queries, samples, meta=None):
    """videos / queries: lists of float arrays [L, D] (already prepared); samples: list of (video index, query index)."""
    dv = int(videos[0].shape[1])
    dt = int(queries[0].shape[1])
    vid_off = np.zeros(len(videos) + 1, dtype=np.int64)
    txt_off = np.zeros(len(queries) + 1, dtype=np.int64)
    for i, v in enumerate(videos):
        if v.shape[1] != dv:
            raise ValueError("all videos must have the same feature width")
        vid_off[i + 1] = vid_off[i] + len(v)
    for i, q in enumerate(queries):
        if q.shape[1] != dt:
            raise ValueError("all queries must have the same feature width")
        txt_off[i + 1] = txt_off[i] + len(q)
    smp = np.asarray(samples, dtype=np.int32).reshape(-1, 2)
    if len(smp) and (smp[:, 0].max() >= len(videos) or smp[:, 1].max() >= len(queries) or smp.min() < 0):
        raise ValueError("sample table refers to a missing video / query")
    arrays = {"vid_off": vid_off, "txt_off": txt_off, "samples": smp}
    sizes = {"vid": int(vid_off[-1]) * dv * 2, "txt": int(txt_off[-1]) * dt * 2}
    header = {"v_feat_dim": dv, "t_feat_dim": dt, "n_videos": len(videos), "n_queries": len(queries), "n_samples": int(len(smp)),
              "dtype": "float16", "meta": meta or {}, "arrays": {}}
    pos = 0
    for name in ("vid_off", "txt_off", "samples", "vid", "txt"):
        nbytes = arrays[name].nbytes if name in arrays else sizes[name]
        header["arrays"][name] = {"offset": pos, "nbytes": int(nbytes)}
        pos = (pos + nbytes + 63) // 64 * 64
    hjson = json.dumps(header).encode()
    hbytes = (len(hjson) + 16 + 4095) // 4096 * 4096 - 16
    hjson = hjson + b" " * (hbytes - len(hjson))
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(hbytes).tobytes())
        f.write(hjson)
        base = f.tell()
        for name in ("vid_off", "txt_off", "samples"):
            f.seek(base + header["arrays"][name]["offset"])
            f.write(arrays[name].tobytes())
        f.seek(base + header["arrays"]["vid"]["offset"])
        for v in videos:
            f.write(np.ascontiguousarray(v, dtype=np.float16).tobytes())
        f.seek(base + header["arrays"]["txt"]["offset"])
        for q in queries:
            f.write(np.ascontiguousarray(q, dtype=np.float16).tobytes())
        f.truncate(base + pos)
    return header


def pack_from_npz_dirs(path, annotations, v_feat_dirs, q_feat_dir, q_feat_type="last_hidden_state", normalize_v=True, normalize_t=True,
                       use_tef=True, max_v_l=None):
    """Build a shard from the reference's on-disk layout: annotations = iterable of dicts with 'qid' and 'vid' (the jsonl lines
    main/dataset.py:80-87 loads); one {vid}.npz['features'] per feature directory, one {qid}.npz[q_feat_type] per query."""
    vids, vid_index, queries, samples, ids = [], {}, [], [], []
    for ann in annotations:
        vid, qid = ann["vid"], ann["qid"]
        if vid not in vid_index:
            feats = [np.load(os.path.join(d, f"{vid}.npz"))["features"] for d in v_feat_dirs]
            vid_index[vid] = len(vids)
            vids.append(prepare_video(feats, normalize_v, use_tef, max_v_l))
        queries.append(prepare_query(np.load(os.path.join(q_feat_dir, f"{qid}.npz"))[q_feat_type], normalize_t))
        samples.append((vid_index[vid], len(queries) - 1))
        ids.append({"qid": qid, "vid": vid})
    return write_shard(path, vids, queries, samples, meta={"ids": ids, "use_tef": bool(use_tef)})


class Shard:
    """Memory-mapped view of a shard file."""

    def __init__(self, path):
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path}: not a univtg_b200 feature shard")
            hbytes = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            self.header = json.loads(f.read(hbytes).decode())
            base = f.tell()
        self.path = path
        a = self.header["arrays"]
        # mode "c" = private copy-on-write mapping: never written by this module, but a writable mapping is what cudaHostRegister
        # accepts (ShardLoader's direct mode page-locks it; read-only mappings need cudaHostRegisterReadOnly, which this platform
        # refuses).  Page-locking a private mapping makes the kernel materialise its pages: direct mode wants shards that fit in RAM.
        mm = np.memmap(path, dtype=np.uint8, mode="c")

        def arr(name, dtype, shape):
            o = base + a[name]["offset"]
            return mm[o:o + a[name]["nbytes"]].view(dtype).reshape(shape)

        self.v_feat_dim, self.t_feat_dim = int(self.header["v_feat_dim"]), int(self.header["t_feat_dim"])
        self.vid_off = np.array(arr("vid_off", np.int64, (-1,)))
        self.txt_off = np.array(arr("txt_off", np.int64, (-1,)))
        self.samples = np.array(arr("samples", np.int32, (-1, 2)))
        self.vid = arr("vid", np.float16, (-1, self.v_feat_dim))
        self.txt = arr("txt", np.float16, (-1, self.t_feat_dim))

    def __len__(self):
        return len(self.samples)

    def video(self, i):
        return self.vid[self.vid_off[i]:self.vid_off[i + 1]]

    def query(self, i):
        return self.txt[self.txt_off[i]:self.txt_off[i + 1]]

    def lengths(self, idx):
        v, q = self.samples[idx, 0], self.samples[idx, 1]
        return self.vid_off[v + 1] - self.vid_off[v], self.txt_off[q + 1] - self.txt_off[q]


def page_extents(regions, page=4096):
    """Page-rounded [lo, hi) extents covering `regions` = [(address, nbytes), ...], overlapping or touching extents merged (a page
    can be registered with the driver only once)."""
    spans = sorted((b // page * page, (b + n + page - 1) // page * page) for b, n in regions if n)
    merged = []
    for lo, hi in spans:
        if merged and lo <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], hi)
        else:
            merged.append([lo, hi])
    return merged


class ShardLoader:
    """Batches of model inputs straight from a shard: src_vid [B, Lv, Dv] / src_txt [B, Lt, Dt] as fp16, masks as float32 (1 = valid),
    zero right-padded to the batch maximum like start_end_collate_mr.  `slots` pinned staging buffers and device buffers are reused
    round-robin; a background thread assembles batch i+1 (per-sample memcpys on `workers` threads) and issues its H2D copy on a side
    stream while the consumer computes on batch i.  The yielded tensors stay valid until `slots - 1` further batches were fetched.

    device=None keeps everything on the host (pinned tensors when CUDA is available) - used by the CPU tests."""

    def __init__(self, shard, batch_size, device=None, shuffle=False, seed=0, rank=0, world=1, drop_last=False, slots=3, workers=4,
                 max_v_l=None, max_q_l=None, direct=True):
        self.shard = shard if isinstance(shard, Shard) else Shard(shard)
        self.batch_size, self.device = int(batch_size), (torch.device(device) if device is not None else None)
        self.shuffle, self.seed, self.rank, self.world, self.drop_last = bool(shuffle), int(seed), int(rank), int(world), bool(drop_last)
        self.slots, self.workers = max(2, int(slots)), max(1, int(workers))
        self.max_v_l, self.max_q_l = max_v_l, max_q_l
        self.epoch = 0
        self._host, self._dev = [None] * self.slots, [None] * self.slots
        self._h2d_done = [None] * self.slots   # the slot's pinned buffers may be rewritten once its last H2D copy has finished
        self._consumed = [None] * self.slots   # the slot's device buffers may be overwritten once the consumer's stream got here
        self._stream = torch.cuda.Stream(device=self.device) if self.device is not None and self.device.type == "cuda" else None
        # direct=True: page-lock the shard's mapping once and let the copy engines pull every sample's rows straight from the page
        # cache (univtg_h2d_gather_batch) - no CPU staging copy.  Falls back to the staged path when the driver refuses.
        self.direct = False
        if direct and self._stream is not None:
            from . import _lib
            lib = _lib.load_library()
            sh = self.shard
            # page-rounded extents of the two feature arrays; they are neighbours in the file, so their end / start pages usually
            # coincide and the extents are merged (a page can be registered once)
            merged = page_extents([(sh.vid.ctypes.data, sh.vid.nbytes), (sh.txt.ctypes.data, sh.txt.nbytes)])
            done = []
            self.direct_error = None
            with torch.cuda.device(self.device):
                for lo, hi in merged:
                    if lib.univtg_host_register(lo, hi - lo, 1) != 0:
                        self.direct_error = lib.univtg_last_error().decode()
                        for b_, _ in done:
                            lib.univtg_host_register(b_, 0, 0)
                        done = None
                        break
                    done.append((lo, hi - lo))
            self._registered = done
            self.direct = done is not None

    def close(self):
        """Undo the page-locking of the shard mapping (direct mode)."""
        if getattr(self, "_registered", None):
            from . import _lib
            lib = _lib.load_library()
            for base, _ in self._registered:
                lib.univtg_host_register(base, 0, 0)
            self._registered = None
            self.direct = False

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def _order(self):
        n = len(self.shard)
        idx = np.arange(n)
        if self.shuffle:
            idx = np.random.default_rng(self.seed + self.epoch).permutation(n)
        idx = idx[self.rank::self.world]  # DistributedSampler-style strided split (main/train_vlp_ddp.py:105-113)
        nb = len(idx) // self.batch_size if self.drop_last else (len(idx) + self.batch_size - 1) // self.batch_size
        return [idx[i * self.batch_size:(i + 1) * self.batch_size] for i in range(nb)]

    def __len__(self):
        return len(self._order())

    def _buffers(self, slot, B, Lv, Lt):
        need = (B, Lv, Lt)
        h = self._host[slot]
        if h is None or h["cap"][0] < B or h["cap"][1] < Lv or h["cap"][2] < Lt:
            cap = (max(B, h["cap"][0] if h else 0), max(Lv, h["cap"][1] if h else 0), max(Lt, h["cap"][2] if h else 0))
            pin = torch.cuda.is_available()
            mk = lambda *shape, dtype: (torch.empty(*shape, dtype=dtype).pin_memory() if pin else torch.empty(*shape, dtype=dtype))  # noqa: E731
            h = {"cap": cap, "vid": mk(cap[0] * cap[1] * self.shard.v_feat_dim, dtype=torch.float16),
                 "txt": mk(cap[0] * cap[2] * self.shard.t_feat_dim, dtype=torch.float16),
                 "vmask": mk(cap[0] * cap[1], dtype=torch.float32), "tmask": mk(cap[0] * cap[2], dtype=torch.float32)}
            self._host[slot] = h
            if self._stream is not None:
                self._dev[slot] = {k: torch.empty_like(v, device=self.device) for k, v in h.items() if k != "cap"}
        return h, need

    def _assemble(self, slot, idx):
        sh = self.shard
        lv, lt = sh.lengths(idx)
        if self.max_v_l is not None:
            lv = np.minimum(lv, self.max_v_l)
        if self.max_q_l is not None:
            lt = np.minimum(lt, self.max_q_l)
        B, Lv, Lt = len(idx), int(lv.max()), int(lt.max())
        if self._h2d_done[slot] is not None:
            self._h2d_done[slot].synchronize()
        h, _ = self._buffers(slot, B, Lv, Lt)
        Dv, Dt = sh.v_feat_dim, sh.t_feat_dim
        if self.direct:
            return self._assemble_direct(slot, idx, lv, lt, B, Lv, Lt)
        vid = h["vid"][:B * Lv * Dv].view(B, Lv, Dv)
        txt = h["txt"][:B * Lt * Dt].view(B, Lt, Dt)
        vmask = h["vmask"][:B * Lv].view(B, Lv)
        tmask = h["tmask"][:B * Lt].view(B, Lt)
        # the gather itself is native (csrc/hostio.cu: one memcpy per sample and modality on a persistent thread pool, GIL released
        # by ctypes) - per-sample numpy statements cost more interpreter time than the 13 MB they move
        smp = sh.samples[idx]
        vrow0 = np.ascontiguousarray(sh.vid_off[smp[:, 0]], dtype=np.int64)
        trow0 = np.ascontiguousarray(sh.txt_off[smp[:, 1]], dtype=np.int64)
        vlen = np.ascontiguousarray(lv, dtype=np.int32)
        tlen = np.ascontiguousarray(lt, dtype=np.int32)
        from . import _lib
        lib = _lib.load_library()
        _lib.check(lib.univtg_host_assemble_batch(vid.data_ptr(), txt.data_ptr(), vmask.data_ptr(), tmask.data_ptr(),
                                                  sh.vid.ctypes.data, sh.txt.ctypes.data, vrow0.ctypes.data, trow0.ctypes.data,
                                                  vlen.ctypes.data, tlen.ctypes.data, B, Lv, Lt, Dv, Dt, self.workers),
                   "univtg_host_assemble_batch")
        out = {"src_vid": vid, "src_vid_mask": vmask, "src_txt": txt, "src_txt_mask": tmask}
        if self._stream is None:
            return out, None
        d = self._dev[slot]
        dev_out = {"src_vid": d["vid"][:B * Lv * Dv].view(B, Lv, Dv), "src_vid_mask": d["vmask"][:B * Lv].view(B, Lv),
                   "src_txt": d["txt"][:B * Lt * Dt].view(B, Lt, Dt), "src_txt_mask": d["tmask"][:B * Lt].view(B, Lt)}
        with torch.cuda.stream(self._stream):
            if self._consumed[slot] is not None:
                self._stream.wait_event(self._consumed[slot])  # kernels still reading this slot's previous batch
            for k in out:
                dev_out[k].copy_(out[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._h2d_done[slot] = ev
        return dev_out, ev

    def _assemble_direct(self, slot, idx, lv, lt, B, Lv, Lt):
        from . import _lib
        lib = _lib.load_library()
        sh = self.shard
        Dv, Dt = sh.v_feat_dim, sh.t_feat_dim
        smp = sh.samples[idx]
        vrow0 = np.ascontiguousarray(sh.vid_off[smp[:, 0]], dtype=np.int64)
        trow0 = np.ascontiguousarray(sh.txt_off[smp[:, 1]], dtype=np.int64)
        vlen = np.ascontiguousarray(lv, dtype=np.int32)
        tlen = np.ascontiguousarray(lt, dtype=np.int32)
        d = self._dev[slot]
        h = self._host[slot]
        dev_out = {"src_vid": d["vid"][:B * Lv * Dv].view(B, Lv, Dv), "src_vid_mask": d["vmask"][:B * Lv].view(B, Lv),
                   "src_txt": d["txt"][:B * Lt * Dt].view(B, Lt, Dt), "src_txt_mask": d["tmask"][:B * Lt].view(B, Lt)}
        # mask staging: the pinned vmask / tmask buffers of the slot, laid out back to back in one scratch tensor
        stage = h.setdefault("mask_stage", torch.empty(h["cap"][0] * (h["cap"][1] + h["cap"][2]), dtype=torch.float32).pin_memory())
        with torch.cuda.device(self.device), torch.cuda.stream(self._stream):
            if self._consumed[slot] is not None:
                self._stream.wait_event(self._consumed[slot])
            _lib.check(lib.univtg_h2d_gather_batch(dev_out["src_vid"].data_ptr(), dev_out["src_txt"].data_ptr(),
                                                   dev_out["src_vid_mask"].data_ptr(), dev_out["src_txt_mask"].data_ptr(), stage.data_ptr(),
                                                   sh.vid.ctypes.data, sh.txt.ctypes.data, vrow0.ctypes.data, trow0.ctypes.data,
                                                   vlen.ctypes.data, tlen.ctypes.data, B, Lv, Lt, Dv, Dt,
                                                   self._stream.cuda_stream), "univtg_h2d_gather_batch")
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._h2d_done[slot] = ev
        return dev_out, ev

    def __iter__(self):
        batches = self._order()
        if not batches:
            return
        results = {}
        lock = threading.Condition()
        depth = self.slots - 1

        def producer():
            for i, idx in enumerate(batches):
                with lock:
                    while i - producer.consumed >= depth:  # never overwrite a slot the consumer may still read
                        lock.wait()
                res = self._assemble(i % self.slots, idx)
                with lock:
                    results[i] = (res, idx)
                    lock.notify_all()

        producer.consumed = 0
        th = threading.Thread(target=producer, daemon=True)
        th.start()
        for i in range(len(batches)):
            with lock:
                while i not in results:
                    lock.wait()
                (tensors, ev), idx = results.pop(i)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
            yield tensors, idx
            if ev is not None:  # everything the consumer enqueued on its stream for this batch precedes the slot's next refill
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream())
                self._consumed[i % self.slots] = done
            with lock:
                producer.consumed = i  # batch i - 1's slot may be recycled once batch i has been handed out
                lock.notify_all()
        th.join()

    def h2d_bytes(self, B, Lv, Lt):
        return B * (Lv * self.shard.v_feat_dim + Lt * self.shard.t_feat_dim) * 2 + B * (Lv + Lt) * 4
