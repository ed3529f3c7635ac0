"""CPU oracle for the Aligner / MAS hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product package
(isp-tts_b200/, importable as isp_tts_b200) never does.
"""
