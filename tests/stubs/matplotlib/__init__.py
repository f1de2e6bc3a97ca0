"""TEST STUB: the reference's eval_batch.py imports matplotlib (eval_batch.py:9), which this image does not have.  The stub
lets the UNMODIFIED script run in tests/test_reference_scripts.py; every plotting call is a no-op."""
