"""Drop-in mirror of the part of the reference's ``mvn`` package that the lifting path touches
(ContextPose/mvn/models/{conpose,pose_hrnet,pose_dformer,networks}.py, mvn/utils/cfg.py)."""
