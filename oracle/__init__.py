"""CPU oracle for the concept-embedding similarity scan.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multimodal_concept_learning_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only as
the checker or the timed CPU baseline -- never as the product path.

Parity status: the reference (AskSid/multimodal_concept_learning) has no tests and
no golden vectors of its own, so the oracle is pinned two ways:

* against the *reference's own functions* imported from ``/root/reference`` in the
  build container (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``, committed), and
* against the third-party primitives the reference calls (scikit-learn
  ``cosine_similarity``, HF ``ForCausalLMLoss``, ``nn.CrossEntropyLoss``,
  ``torch.topk`` / ``argmax``), which are importable wherever the tests run.
"""
