"""Import-path shim: the reference's ``src`` package backed by probabilisticdeepdiffusionmodels_b200."""
