"""Stand-in for ``rdkit.Chem.inchi`` (rdkit is not installed in this image).

The reference only takes ``logger`` from it (utilities.py:11, used at :1544,:1551).
"""
import logging

logger = logging.getLogger("rdkit.stub")
