"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see bbme_oracle.c).  Imported by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs; never by blockbasedmotionestimation_b200."""
