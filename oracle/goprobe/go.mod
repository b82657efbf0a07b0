module kvidx/goprobe

go 1.22

require github.com/fxamacker/cbor/v2 v2.7.0
