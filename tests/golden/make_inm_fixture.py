"""tests/golden/testInmLoader.out: the INM mesh file the reference's own loader test reads
(/root/reference/meshes/testInmLoader.out, src/test/sequence/TestInmMeshLoader.cpp:19-68), copied byte for byte so that the
test travels to machines without the reference tree.  Run where /root/reference exists: python tests/golden/make_inm_fixture.py"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
shutil.copyfile("/root/reference/meshes/testInmLoader.out", os.path.join(HERE, "testInmLoader.out"))
os.chmod(os.path.join(HERE, "testInmLoader.out"), 0o644)
print(open(os.path.join(HERE, "testInmLoader.out")).read())
