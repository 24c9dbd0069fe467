!  Box size of the drop-in executable.  pymulskips rewrites the LenX line below before every compilation
!  (pymulskips/setuprun.py:499-522, the same edit it makes in mulskips-source/modules/defsystem.f:27); the B200 host
!  program reads it at START-UP from <directory of mulskips.e>/modules/defsystem.f, so the box needs no recompilation
!  (`--box X Y Z` or MSKC_BOX override it).
       INTEGER, PARAMETER :: LenX=480, LenY=480, LenZ=480
