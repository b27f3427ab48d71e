"""CPU oracle for the LSH-attention / reversible / chunked-FFN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``reformer_tts_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or as
the timed CPU baseline - never as the product path.

Parity status (see DESIGN.md "Oracle"):

* ``lsh_hf``  - restatement of ``transformers`` ``LSHSelfAttention`` (5.5.0 source,
  the only LSH implementation physically present in the build container).  PINNED:
  checked stage by stage against the real class (``tests/golden/make_golden.py``
  ran it here; vectors committed under ``tests/golden/``).
* ``lsh_rp``  - restatement of ``reformer-pytorch==0.19.1`` (pinned at
  ref:requirements.txt:11).  That package is NOT in /root/reference and cannot be
  installed (no network), and the reference has no tests or golden vectors for
  the path, so this half is **parity unpinned** against the package itself.  It is
  pinned only by closed-form known-answer tests (dense-attention equivalence,
  stable-sort identity, round-merge identities) and by sharing every primitive
  with ``lsh_hf`` (hash, sort, look-one-back, merge), which *is* pinned.
* ``reversible`` - restatement of ref:reformer_tts/model/reversible.py; PINNED
  against the reference file itself (imported in the build container; gradients
  committed as golden vectors).
"""
