__version__ = "0.1.11+b200.1"  # API level of the reference fork (gsplat/version.py:1) + this backend
