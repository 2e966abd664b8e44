"""Drop-in ``model`` package: same import paths as the reference (``model.point_net2``,
``model.project_to_2d``), backed by libsn2_b200.so.  Put this directory's parent
(stratanet2-vegetation-coverage-maps_b200/) on sys.path in place of the reference repo root."""
