// tu.cu -- empty translation unit: the Makefile force-includes one reference test header into it
