"""CPU oracle for the ZPAQ block codec path -- TEST INFRASTRUCTURE ONLY.
Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from
zpaqsharp_b200/."""
