"""oracle/ — TEST INFRASTRUCTURE: the CPU restatement of the reference's preqx timestep (oracle.c) and the
reference's own PPM twin (oracle/_ref). Only tests/, __graft_entry__.smoke() and bench.py's CPU arm import this."""
