"""CPU oracle for the torchctr embedding / interaction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``torchctr_b200`` imports this package.
The only permitted callers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker or the reported CPU baseline, never as the product path.

What it restates (reference paths relative to ``/root/reference``):

* ``oracle.hashing``    -- ``torchctr/utils.py:103-119`` (``hash_bucket``) on top of
  scikit-learn's ``murmurhash3_32`` (MurmurHash3_x86_32; third-party, pinned
  ``scikit-learn>=1.5.1`` in ``setup.py:33``, 1.9.0 installed when the golden
  vectors were generated).  A C restatement lives in ``oracle/murmur3.c``.
* ``oracle.vocab``      -- ``torchctr/transformer.py:451-498`` (vocab build / grow / lookup).
* ``oracle.embedding``  -- ``torchctr/models/dnn.py:53-59`` (masked gather + sum pool),
  its autograd (dense ``[V, D]`` gradient) and ``torchctr/nn/embedding.py:63-95``
  (``DynamicEmbedding`` growth and state-dict merge).
* ``oracle.optim``      -- the optimizers the reference Trainer drives
  (``torchctr/trainer.py:303``): torch Adagrad / SparseAdam / SGD restated on
  unique rows, plus row-wise Adagrad (no torch equivalent).
* ``oracle.models``     -- ``torchctr/models/dnn.py:10-82`` (``DNN``) and the
  DeepFM / DCN-v2 definitions the reference lacks.

Pinning: the restatements are checked in ``tests/test_oracle_golden.py`` against
fixtures under ``tests/golden/`` that were produced by importing the real
reference (and scikit-learn) in the build container with
``tests/golden/make_golden.py``.  FM / DeepFM / DCN-v2 / mean pooling / row-wise
Adagrad do not exist in the reference: for those the header of the respective
function says "parity unpinned by the reference".
"""
