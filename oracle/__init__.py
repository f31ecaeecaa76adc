"""CPU oracle package — TEST INFRASTRUCTURE, NOT PRODUCT (see rt_oracle.h)."""
