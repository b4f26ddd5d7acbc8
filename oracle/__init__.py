"""Test infrastructure: CPU oracle for the SSHash lookup path.

`oracle.port`  -- plain-C restatement (oracle/sshash_oracle.c -> liboracle.so), travels everywhere.
`oracle.ref`   -- the unmodified reference compiled from /root/reference (oracle/_ref/*.so).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (sshash_b200/) never does.
"""
