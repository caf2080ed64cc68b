"""N1 bias library surface (reference: plancklens/n1/).  See n1.py."""
