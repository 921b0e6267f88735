"""CPU oracle for the CLIP-PPO observation path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / CPU baseline.

Parity status: the reference ships no golden vectors.  The pins are
outputs of the reference's *own* functions, generated in the build container by
``oracle/make_goldens.py`` (which imports ``/root/reference``) and committed under
``tests/golden/``.  The ViT tower lives in the un-vendored dependency openai/CLIP
(``requirements.txt:12``, unpinned HEAD); its restatement in ``oracle/vit.py`` is pinned
against ``transformers.CLIPVisionModelWithProjection`` with copied random weights.
"""
