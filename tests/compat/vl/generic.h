/* harness compatibility pack: test/test.cpp includes VLFeat headers it never uses */
