"""Device-side replacements for the reference's per-sample Python data path and evaluators
(SURVEY.md §8(f) ranks 3 and 4).

``LayoutStore`` holds a dataset's layouts in CSR form on the GPU (built once from the JSON schema
the reference reads, src/modelling/datasets.py:35-37); ``build_batch`` replaces
``StltDataset.__getitem__`` + ``StltCollater.__call__`` (datasets.py:52-125,243-288) with one
kernel and returns the same batch dict, bit for bit. ``TopKCounter`` replaces the per-batch
``.cpu()`` bookkeeping of ``EvaluatorSomething`` (src/utils/evaluation.py:21-34).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import torch

from . import lib as _lib
from .configs import ACTION_GENOME, SOMETHING_ELSE
from .prepare import _prep_handle

# category2id of the two datasets (reference src/modelling/configs.py:30-78)
_AG_NAMES = ["pad", "cls", "chair", "book", "medicine", "vacuum", "food", "groceries", "floor", "mirror",
             "closet/cabinet", "doorway", "paper/notebook", "picture", "phone/camera", "sofa/couch", "sandwich",
             "cup/glass/bottle", "towel", "box", "blanket", "television", "bag", "refrigerator", "table", "light",
             "broom", "shoe", "doorknob", "bed", "window", "shelf", "door", "pillow", "laptop", "dish", "clothes",
             "person"]
CATEGORY2ID = {
    "something": {"pad": 0, "hand": 1, "object": 2, "cls": 3},
    "action_genome": {n: i for i, n in enumerate(_AG_NAMES)},
}
_SPECS = {"something": SOMETHING_ELSE, "action_genome": ACTION_GENOME}


class LayoutStore:
    """CSR copy of a layout dataset on one CUDA device."""

    def __init__(self, dataset_name: str, videos: Sequence[dict], videoid2size: Dict[str, Sequence[int]],
                 device="cuda", score_threshold: float = 0.5, layout_num_frames: int = 16):
        if dataset_name not in CATEGORY2ID:
            raise ValueError(f"{dataset_name} does not exist!")
        self.dataset_name = dataset_name
        self.score_threshold = float(score_threshold)
        self.layout_num_frames = int(layout_num_frames)
        c2i = CATEGORY2ID[dataset_name]
        vfo, foo, boxes, cats, scores, sizes = [0], [0], [], [], [], []
        max_objects = -1
        self.video_ids: List[str] = []
        for video in videos:
            self.video_ids.append(video["id"])
            sizes.append(list(videoid2size[video["id"]]))
            for frame in video["frames"]:
                kept = 0
                for e in frame["frame_objects"]:
                    boxes.append([e["x1"], e["y1"], e["x2"], e["y2"]])
                    cats.append(c2i[e["category"]])
                    scores.append(e["score"])
                    kept += e["score"] >= self.score_threshold
                max_objects = max(max_objects, kept)  # datasets.py:38-47
                foo.append(len(cats))
            vfo.append(len(foo) - 1)
        self.max_num_objects = max_objects
        self.num_frames = torch.tensor([vfo[i + 1] - vfo[i] for i in range(len(vfo) - 1)], dtype=torch.int64)
        dev = torch.device(device)
        self.device = dev
        self.video_frame_offsets = torch.tensor(vfo, dtype=torch.int64, device=dev)
        self.frame_object_offsets = torch.tensor(foo, dtype=torch.int64, device=dev)
        self.obj_boxes = torch.tensor(boxes, dtype=torch.float64, device=dev).reshape(-1, 4)
        self.obj_categories = torch.tensor(cats, dtype=torch.int64, device=dev)
        self.obj_scores = torch.tensor(scores, dtype=torch.float64, device=dev)
        self.video_sizes = torch.tensor(sizes, dtype=torch.int64, device=dev).reshape(-1, 2)

    def __len__(self):
        return len(self.video_ids)

    def _c_store(self) -> _lib.StltLayoutStore:
        return _lib.StltLayoutStore(self.video_frame_offsets.data_ptr(), self.frame_object_offsets.data_ptr(),
                                    self.obj_boxes.data_ptr(), self.obj_categories.data_ptr(),
                                    self.obj_scores.data_ptr(), self.video_sizes.data_ptr())

    def build_batch(self, video_index: Sequence[int], frame_indices: Optional[List[List[int]]] = None,
                    check: bool = False) -> Dict[str, torch.Tensor]:
        """Batch dict for videos ``video_index`` exactly as DataLoader(StltDataset, collate_fn=StltCollater)
        yields it (minus ``labels``). ``frame_indices`` overrides the test-time frame sampling (training
        uses random indices, src/utils/data_utils.py:32-44)."""
        spec = _SPECS[self.dataset_name]
        ft = spec["frame_types"]
        dev = self.device
        B, T, S = len(video_index), self.layout_num_frames, self.max_num_objects + 1
        vid_cpu = torch.as_tensor(list(video_index), dtype=torch.int64)
        if frame_indices is None:
            n_s = torch.clamp(self.num_frames[vid_cpu], max=T) if B else torch.zeros(0, dtype=torch.int64)
            fi_dev = ns_dev = None
        else:
            n_s = torch.tensor([len(f) for f in frame_indices], dtype=torch.int64)
            fi = torch.zeros((B, T), dtype=torch.int64)
            for b, f in enumerate(frame_indices):
                fi[b, : len(f)] = torch.as_tensor(f, dtype=torch.int64)
            fi_dev, ns_dev = fi.to(dev), n_s.to(dev)
        L = int(n_s.max()) + 1 if B else 1
        out = {
            "video_id": [self.video_ids[i] for i in vid_cpu.tolist()],
            "categories": torch.empty((B, L, S), dtype=torch.int64, device=dev),
            "boxes": torch.empty((B, L, S, 4), dtype=torch.float32, device=dev),
            "frame_types": torch.empty((B, L), dtype=torch.int64, device=dev),
            "lengths": torch.empty((B,), dtype=torch.int64, device=dev),
            "src_key_padding_mask_boxes": torch.empty((B, L, S), dtype=torch.bool, device=dev),
            "src_key_padding_mask_frames": torch.empty((B, L), dtype=torch.bool, device=dev),
        }
        if spec["scores"]:  # the collater keeps scores only for Action Genome (datasets.py:253-260)
            out["scores"] = torch.empty((B, L, S), dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        vid_dev = vid_cpu.to(dev)
        ids = _lib.StltLayoutIds(spec["cls_id"], ft["pad"], ft["regular"], ft["empty"], ft["extract"])
        store = self._c_store()
        lib = _lib.load_library()
        with torch.cuda.device(dev):
            handle = _prep_handle(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.stlt_build_batch(
                handle, stream, ctypes.byref(store), ctypes.byref(ids), vid_dev.data_ptr(),
                fi_dev.data_ptr() if fi_dev is not None else None, ns_dev.data_ptr() if ns_dev is not None else None,
                B, T, L, S, self.score_threshold, out["categories"].data_ptr(), out["boxes"].data_ptr(),
                out["scores"].data_ptr() if "scores" in out else None, out["frame_types"].data_ptr(),
                out["lengths"].data_ptr(), out["src_key_padding_mask_boxes"].data_ptr(),
                out["src_key_padding_mask_frames"].data_ptr(), status.data_ptr())
            _lib.check(handle, rc)
        if check and int(status.item()) != 0:
            raise RuntimeError(f"stlt_build_batch reported inconsistent layout data (code {int(status.item())})")
        return out


class TopKCounter:
    """Device-resident top-1 / top-5 hit counters (EvaluatorSomething.process without the per-batch sync)."""

    def __init__(self, total_instances: int, device="cuda"):
        self.total_instances = total_instances
        self.device = torch.device(device)
        self.counters = torch.zeros(2, dtype=torch.int64, device=self.device)

    def reset(self):
        self.counters.zero_()

    def process(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        if logits.dtype != torch.float32 or labels.dtype != torch.int64 or logits.device.type != "cuda":
            raise TypeError("logits must be CUDA float32 [rows, classes], labels int64 [rows]")
        logits, labels = logits.contiguous(), labels.contiguous()
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            handle = _prep_handle(self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(handle, lib.stlt_topk_count(handle, stream, logits.data_ptr(), labels.data_ptr(),
                                                   logits.shape[0], logits.shape[1], self.counters.data_ptr()))

    def evaluate(self, name: str = "stlt") -> Dict[str, float]:
        top1, top5 = self.counters.tolist()  # the only synchronisation
        return {f"{name}_top1_accuracy": top1 / self.total_instances,
                f"{name}_top5_accuracy": top5 / self.total_instances}


class CharadesMapEvaluator:
    """Device-resident EvaluatorActionGenome (reference src/utils/evaluation.py:61-98): ``process`` appends
    sigmoid(logits) and the multi-hot labels to device buffers without synchronising, ``evaluate`` runs the
    Charades mAP (evaluation.py:100-132) on the device and reads back one number."""

    def __init__(self, total_instances: int, total_classes: int, logit_names=("stlt",), device="cuda"):
        self.total_instances, self.total_classes, self.logit_names = total_instances, total_classes, logit_names
        self.device = torch.device(device)
        self.predictions = torch.zeros((total_instances, total_classes), dtype=torch.float32, device=self.device)
        self.ground_truths = torch.zeros((total_instances, total_classes), dtype=torch.float32, device=self.device)
        self.best_mean_average_precision = 0.0
        self.index = 0

    def reset(self):
        self.index = 0
        self.predictions.zero_()
        self.ground_truths.zero_()

    def process(self, logits, labels: torch.Tensor) -> None:
        x = logits["stlt"] if isinstance(logits, dict) else logits  # Action Genome only for STLT (evaluation.py:77)
        if x.dtype != torch.float32 or x.device.type != "cuda":
            raise TypeError("logits must be CUDA float32 [rows, classes]")
        size = x.shape[0]
        if self.index + size > self.total_instances:
            raise ValueError("more rows than total_instances")
        x = x.contiguous()
        y = labels.to(device=self.device, dtype=torch.float32).contiguous()
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            handle = _prep_handle(self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(handle, lib.stlt_map_accumulate(handle, stream, x.data_ptr(), y.data_ptr(), size,
                                                       self.total_classes, self.predictions[self.index].data_ptr(),
                                                       self.ground_truths[self.index].data_ptr()))
        self.index += size

    def evaluate(self) -> Dict[str, float]:
        ap = torch.empty(self.total_classes, dtype=torch.float64, device=self.device)
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            handle = _prep_handle(self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(handle, lib.stlt_charades_map(handle, stream, self.predictions.data_ptr(),
                                                     self.ground_truths.data_ptr(), self.total_instances,
                                                     self.total_classes, ap.data_ptr(), out.data_ptr()))
        self.average_precisions = ap
        return {"map": float(out.item())}  # the only synchronisation

    def is_best(self) -> bool:
        metrics = self.evaluate()
        if metrics["map"] > self.best_mean_average_precision:
            self.best_mean_average_precision = metrics["map"]
            return True
        return False
