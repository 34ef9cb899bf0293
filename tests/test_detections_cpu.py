"""Detection wire format (class_query_vad_b200/detections.py): binary dump round trip and the reference's `{rank}.txt` text
format (utils/video_action_recognition.py:231-236), checked against the committed PostProcessAVA fixture of the reference."""
import os
import numpy as np
import torch

from class_query_vad_b200.detections import save_detections, load_detections, write_reference_text, read_reference_text

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fixture():
    z = np.load(os.path.join(GOLD, "criterion_ava.npz"))
    det = z["det"]                                   # reference PostProcessAVA output, [B, nq, K + 5] = [scores | boxes | person]
    ids = [f"vid{b:03d},{900 + b:04d}" for b in range(det.shape[0])]          # AVA frame ids contain a comma (ava_frame.py)
    return ids, det, z["pred_logits"].shape[-1]


def test_binary_round_trip(tmp_path):
    ids, det, K = _fixture()
    p = tmp_path / "det.npz"
    save_detections(p, ids, torch.from_numpy(det), K)
    ids2, det2 = load_detections(p)
    assert ids2 == ids and np.array_equal(det2, det)


def test_reference_text_format(tmp_path):
    ids, det, K = _fixture()
    p = tmp_path / "0.txt"
    write_reference_text(p, ids, det, K)
    # the lines the reference's loop would have written from its own buffers (buff_anno = boxes, buff_output = scores, buff_binary)
    nq = det.shape[1]
    expect = []
    for b in range(det.shape[0]):
        for q in range(nq):
            data = np.concatenate([det[b, q, K:K + 4], det[b, q, :K], det[b, q, K + 4:]])
            expect.append("{} {}\n".format(ids[b], data.tolist()))
    assert open(p).readlines() == expect
    lid, boxes, scores, person = read_reference_text(p, K)
    assert lid == [i for i in ids for _ in range(nq)]
    flat = det.reshape(-1, K + 5).astype(np.float64)
    assert np.array_equal(boxes, flat[:, K:K + 4]) and np.array_equal(scores, flat[:, :K]) and np.array_equal(person, flat[:, K + 4])


def test_shape_errors(tmp_path):
    ids, det, K = _fixture()
    import pytest
    with pytest.raises(ValueError):
        save_detections(tmp_path / "x.npz", ids[:-1], det, K)
    with pytest.raises(ValueError):
        write_reference_text(tmp_path / "x.txt", ids, det[..., :-1], K)
