"""Test infrastructure: the .rda / .RDS reader lives in the product (pareben_b200/rdata.py, the wire format of the
reference's bundled inputs); re-exported here for the golden-fixture generators and tests."""
from pareben_b200.rdata import *  # noqa: F401,F403
from pareben_b200.rdata import read_rda, read_rds  # noqa: F401
