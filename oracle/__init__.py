"""ORACLE package: CPU restatements of the reference hot path.  Test infrastructure;
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
