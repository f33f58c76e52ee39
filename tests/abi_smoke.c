#include <stdio.h>
#include "vidi_b200.h"
int main(void) {
    printf("abi=%d err=[%s]\n", vidi_abi_version(), vidi_last_error());
    return vidi_abi_version() > 0 ? 0 : 1;
}
