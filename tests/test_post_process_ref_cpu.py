"""Oracle.post_process against outputs of the REFERENCE's own `post_process` (groma/train/train_det.py:97-131), executed from
the reference source by tests/golden/make_post_process_golden.py and committed as a fixture."""
import importlib.util
import os

import torch

from oracle.groma_oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_post_process_golden", os.path.join(HERE, "golden", "make_post_process_golden.py"))
gold = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gold)


def _same(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g["labels"], w["labels"])
        assert torch.equal(g["scores"], w["scores"])
        assert torch.allclose(g["boxes"], w["boxes"], rtol=0, atol=1e-4)     # absolute pixels; fp32 op order may differ by an ulp


def test_oracle_post_process_equals_reference_fixture():
    fx = torch.load(os.path.join(HERE, "golden", "post_process_ref.pt"))
    for c, want in zip(fx["cases"], fx["results"]):
        _same(Oracle.post_process(c["coco"], c["boxes"], c["sizes"], c["threshold"], c["top_k"]), want)


def test_fixture_is_what_the_reference_source_produces_now():
    if not os.path.exists(gold.REF):
        import pytest
        pytest.skip("/root/reference not present (GPU box)")
    fx = torch.load(os.path.join(HERE, "golden", "post_process_ref.pt"))
    for want, live in zip(fx["results"], gold.run_reference()):
        for w, l in zip(want, live):
            assert all(torch.equal(w[k], l[k]) for k in w)
