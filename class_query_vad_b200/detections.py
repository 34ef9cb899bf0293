"""Detection wire format of the evaluation loop.

The reference's validate_* loops (utils/video_action_recognition.py:141-156, 231-236) copy the post-processed scores / boxes /
person probabilities of every clip to the host, keep them in Python lists, and each rank writes a text file `{rank}.txt` with one
line per (clip, query):  "<frame id> [x1, y1, x2, y2, s_0 .. s_{K-1}, p_person]";  rank 0 re-parses all files after a barrier
(evaluates/evaluate_ava.py:106-145, evaluate_ucf.py / evaluate_jhmdb.py the same split).

Here the per-clip detections stay one device tensor [B, nq, K + 5] = [scores | box xyxy | p_person] (PostProcessAVA.detections,
cqvad_postprocess_ava), are all-gathered once (dist.gather_detections) and leave the device in ONE copy:

  save_detections(path, frame_ids, det)   binary dump (numpy .npz: ids, boxes, scores, person) -- the native format
  load_detections(path)                   -> (frame_ids, det)
  write_reference_text(path, ...)         the reference's own `{rank}.txt` lines, byte for byte what its loop would write for the
                                          same numbers, so evaluates/*.load_detection_from_path read it unchanged
  read_reference_text(path, K)            the inverse (what the evaluators parse), for tests / migration
"""
import numpy as np
import torch


def _split(det, num_classes):
    det = det.detach().to("cpu", torch.float32).numpy() if isinstance(det, torch.Tensor) else np.asarray(det, dtype=np.float32)
    if det.ndim != 3 or det.shape[-1] != num_classes + 5:
        raise ValueError(f"detections must be [clips, queries, {num_classes + 5}], got {det.shape}")
    return det[..., :num_classes], det[..., num_classes:num_classes + 4], det[..., num_classes + 4:]


def save_detections(path, frame_ids, det, num_classes):
    """Binary dump of gathered detections `det` [clips, nq, K+5] with one frame id per clip."""
    scores, boxes, person = _split(det, num_classes)
    ids = np.asarray([str(i) for i in frame_ids])
    if len(ids) != scores.shape[0]:
        raise ValueError(f"{len(ids)} frame ids for {scores.shape[0]} clips")
    with open(path, "wb") as f:
        np.savez(f, ids=ids, boxes=boxes, scores=scores, person=person)


def load_detections(path):
    """-> (frame_ids list[str], det float32 [clips, nq, K+5] in the [scores | boxes | person] layout)."""
    with np.load(path, allow_pickle=False) as z:
        return [str(i) for i in z["ids"]], np.concatenate([z["scores"], z["boxes"], z["person"]], -1)


def write_reference_text(path, frame_ids, det, num_classes):
    """The reference's per-rank text file (utils/video_action_recognition.py:231-236): one line per (clip, query), the clip's frame
    id repeated for each of its queries (:157-158), numbers printed as the Python floats of the float32 values."""
    scores, boxes, person = _split(det, num_classes)
    with open(path, "w") as f:
        for b, fid in enumerate(frame_ids):
            for q in range(scores.shape[1]):
                data = np.concatenate([boxes[b, q], scores[b, q], person[b, q]])
                f.write("{} {}\n".format(fid, data.tolist()))


def read_reference_text(path, num_classes):
    """Parses `{rank}.txt` the way evaluates/evaluate_ava.py:115-123 does -> (ids per line, boxes [n,4], scores [n,K], person [n])."""
    ids, rows = [], []
    with open(path) as f:
        for line in f:
            ids.append(line.split(" [")[0])
            rows.append([float(x) for x in line.split(" [")[1].split("]")[0].split(",")])
    a = np.asarray(rows, dtype=np.float64).reshape(len(rows), num_classes + 5)
    return ids, a[:, :4], a[:, 4:4 + num_classes], a[:, 4 + num_classes]
