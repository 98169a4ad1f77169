"""Import shim: the reference imports psycopg2 at module level (terminology/mesh.py:8)."""
