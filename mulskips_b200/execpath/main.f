!  Socket port of a laser-annealing run.  pymulskips rewrites the line below when it is given a port
!  (pymulskips/setuprun.py:524-545, the edit it makes in mulskips-source-LaserAnnealing/main.f); the B200 host program
!  takes the port from `-p` on its command line, else from here when the value is positive.
      port = 0
