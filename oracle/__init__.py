"""CPU oracle for the BERT-large pretraining hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline -- never as the thing shipped.  The product path
(``deeplearningexamples_b200``) never imports this package and raises if its
CUDA library is missing.
"""
