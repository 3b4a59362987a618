"""ORACLE — test infrastructure only (see darknet_ref.py).  Never imported by the product package."""
