"""CPU oracle for the Locator hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference algorithm
(kr-colab/locator, ``locator/locator.py``) and of the third-party semantics it
leans on (scikit-allel ingest, numpy legacy RNG, Keras BatchNorm/Dense/Adam and
callbacks).  It exists to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product package ``locator_b200`` never does.

Parity status: the reference ships no tests and its arithmetic lives in
TensorFlow/Keras + scikit-allel, none of which is installable here, so the Keras
model math is "parity unpinned" against a live TF run.  What *is* pinned:
  * ingest / filter / imputation / split / bootstrap / jacknife results against vectors produced by the
    reference's OWN functions, compiled from /root/reference/locator/locator.py and run in the build
    container by tests/golden/make_reference_vectors.py (tests/golden/reference_vectors.*), and
    against facts derived independently from the reference's fixture data (tests/golden/make_golden.py);
  * the model math against torch autograd (tests/test_oracle.py).
"""
